/*
 * refbench.c -- times the REFERENCE runtime (oracle/_ref/libavdspruntime<fmt>.so, compiled from
 * /root/reference by oracle/Makefile with the reference's own flags) on the host cores.
 * TEST / BASELINE INFRASTRUCTURE ONLY: bench.py's cpu_baseline leg and `bench.py --impl reference`.
 *
 * The reference keeps its sample-rate tables and all dither/PRNG state in process globals
 * (runtime/dsp_runtime.c:36-38,103-110, runtime/dsp_tpdf.h:11-13,23,33), so streams run one after
 * another inside a process ("stream-major") and host cores are used with fork(): one worker per core,
 * each with its own copy of the library's globals.  Per stream: private [code|data] buffer,
 * dspRuntimeInit(seed = stream index), then every frame with cores ascending on one io[] -- the same
 * loop nest as linux/avdsp_plugin.c:95-142 / osx/dsprunosx.c:90-91 (canonical order).
 *
 * Synthetic PCM is the documented LCG (avdsp_b200/synth.py): u0 = 0x9E3779B9*(s+1),
 * u <- u*1664525+1013904223 per (frame, channel), sample = (int32)u >> 2.
 *
 * usage: refbench <libavdspruntimeN.so> <program.bin> <fmt> <fs> <workers> <streams_per_worker> <frames> [dither]
 * prints one JSON line: frames, seconds (max over workers), frames_per_s (aggregate), checksum.
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <time.h>
#include <unistd.h>

typedef int (*init_fn)(int32_t *, int, int, int, int);
typedef int32_t *(*find_fn)(int32_t *, int);
typedef int32_t *(*begin_fn)(int32_t *);
typedef int (*run_fn)(int32_t *, int *, void *);

static double now(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }

int main(int argc, char **argv) {
    if (argc < 8) { fprintf(stderr, "usage: %s lib prog.bin fmt fs workers streams_per_worker frames [dither]\n", argv[0]); return 2; }
    const char *libp = argv[1], *progp = argv[2];
    int fmt = atoi(argv[3]), fs = atoi(argv[4]), workers = atoi(argv[5]), spw = atoi(argv[6]), frames = atoi(argv[7]);
    int dither = argc > 8 ? atoi(argv[8]) : 31;
    FILE *f = fopen(progp, "rb");
    if (!f) { perror(progp); return 2; }
    static int32_t prog[65536];
    int nw = (int)fread(prog, 4, 65536, f);
    fclose(f);
    if (nw < 12) { fprintf(stderr, "short program\n"); return 2; }
    int total = prog[1], dsz = prog[2];
    uint32_t inMask = 0, outMask = 0;
    for (int p = 0;;) {                       /* union of the DSP_CORE io bitmaps */
        int sk = prog[p] & 0xFFFF, op = (uint32_t)prog[p] >> 16;
        if (!sk) break;
        if (op == 3) { inMask |= (uint32_t)prog[p + 1]; outMask |= (uint32_t)prog[p + 2]; }
        p += sk;
    }
    if (!inMask && !outMask) { inMask = (uint32_t)prog[9]; outMask = (uint32_t)prog[10]; }
    int inIdx[32], outIdx[32], nIn = 0, nOut = 0;
    for (int k = 0; k < 32; k++) { if (inMask >> k & 1) inIdx[nIn++] = k; if (outMask >> k & 1) outIdx[nOut++] = k; }

    double *shared = mmap(NULL, sizeof(double) * 2 * (size_t)workers, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    for (int wk = 0; wk < workers; wk++) {
        pid_t pid = fork();
        if (pid == 0) {
            void *h = dlopen(libp, RTLD_NOW | RTLD_LOCAL);
            if (!h) { fprintf(stderr, "%s\n", dlerror()); _exit(3); }
            char name[32]; snprintf(name, sizeof name, "dspRuntime_%d", fmt);
            init_fn init = (init_fn)dlsym(h, "dspRuntimeInit");
            find_fn find = (find_fn)dlsym(h, "dspFindCore");
            begin_fn begin = (begin_fn)dlsym(h, "dspFindCoreBegin");
            run_fn run = (run_fn)dlsym(h, name);
            if (!init || !find || !begin || !run) _exit(4);
            int size = total + dsz + 16;
            int32_t *buf = aligned_alloc(64, sizeof(int32_t) * (size_t)(size + 16));
            int32_t *x = malloc(sizeof(int32_t) * (size_t)frames * (nIn ? nIn : 1));
            double busy = 0; uint64_t sum = 0;
            for (int k = 0; k < spw; k++) {
                int s = wk * spw + k;
                memset(buf, 0, sizeof(int32_t) * (size_t)(size + 16));
                memcpy(buf, prog, sizeof(int32_t) * (size_t)total);
                int rc = init(buf, size, fs, s, dither);
                if (rc < 0) _exit(5);
                int32_t *cores[32]; int nc = 0;
                for (int c = 1; c <= 32; c++) {
                    int32_t *p = find(buf, c);
                    if (!p) break;
                    cores[nc++] = begin(p);
                    if (p == buf) break;
                }
                uint32_t u = 0x9E3779B9u * (uint32_t)(s + 1);
                for (long i = 0; i < (long)frames * nIn; i++) { u = u * 1664525u + 1013904223u; x[i] = ((int32_t)u) >> 2; }
                int *data = buf + rc;
                int32_t io[32];
                double t0 = now();
                for (int n = 0; n < frames; n++) {
                    memset(io, 0, sizeof io);
                    for (int c = 0; c < nIn; c++) io[inIdx[c]] = x[(long)n * nIn + c];
                    for (int c = 0; c < nc; c++) run(cores[c], data, io);
                    for (int c = 0; c < nOut; c++) sum += (uint32_t)io[outIdx[c]];
                }
                busy += now() - t0;
            }
            shared[2 * wk] = busy; shared[2 * wk + 1] = (double)(sum & 0xFFFFFFFFu);
            _exit(0);
        }
    }
    int bad = 0;
    for (int wk = 0; wk < workers; wk++) { int st; wait(&st); if (!WIFEXITED(st) || WEXITSTATUS(st)) bad = 1; }
    if (bad) { fprintf(stderr, "a worker failed\n"); return 1; }
    double tmax = 0, cks = 0;
    for (int wk = 0; wk < workers; wk++) { if (shared[2 * wk] > tmax) tmax = shared[2 * wk]; cks += shared[2 * wk + 1]; }
    double nfr = (double)workers * spw * frames;
    printf("{\"frames\": %.0f, \"seconds\": %.6f, \"frames_per_s\": %.1f, \"n_in\": %d, \"n_out\": %d, \"workers\": %d, \"checksum\": %.0f}\n",
           nfr, tmax, nfr / tmax, nIn, nOut, workers, cks);
    return 0;
}
