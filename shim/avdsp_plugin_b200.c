/*
 * avdsp_plugin_b200.c -- the AVDSP ALSA external filter plugin with the per-frame interpreter loop replaced by ONE batched
 * call into libavdsp_b200.so per period.
 *
 * What it replaces, in /root/reference/module_avdsp/linux/avdsp_plugin.c:
 *   dsp_transfer  :71-163   "for core: for frame: gather io[], dspRuntime_<fmt>(), scatter"  ->  avdsp_b200_process_pcm
 *   dsp_init      :172-189  dspRuntimeReset(rate, 0, dither)                                   ->  avdsp_b200_create / _reset
 *   open function :199-370  dspReadBuffer + dspRuntimeInit + core / channel discovery          ->  file read + avdsp_b200_describe
 *                                                                                                  + dspFindCore (same rules)
 * Same asound.conf keys (dspprog, dither 0|7..31, timestat 0..60, tagoutput 0..31, slave; :228-289), same channel
 * convention (inputs io[8..15], outputs io[0..7], channel count = highest used index + 1; :29-32, :337-351), same sample
 * formats (S16 / S24_3LE / S32 in, always S32 out; :109-121, :138, :364).  Two keys are new:
 *   order  "plugin" (default): the reference plugin's core-major loop nest inside each period with a fresh io[] per (core,
 *          frame) -- bit for bit what the old plugin produced, whatever the program;  "canonical": frame-major, cores
 *          ascending (what XMOS targets and osx/dsprunosx.c do), which lets the executor use its fused kernels.  Programs
 *          whose cores do not talk to each other give identical output either way.
 *   device CUDA ordinal (default 0).
 *
 * Builds against real alsa-lib, or against shim/alsa_stub (a 70-line stand-in) where alsa headers are missing.
 * Written from scratch for the batched call; nothing is taken over from the reference's file but the ALSA contract.
 */
#include <alsa/asoundlib.h>
#include <alsa/pcm_external.h>
#include <stdint.h>
#include <time.h>
#include "avdsp_b200.h"

#define AVDSP_IO_SLOTS   16      /* the linux host uses the low 16 io slots (avdsp_plugin.c:28) */
#define AVDSP_OUT_BASE   0
#define AVDSP_IN_BASE    8
#define AVDSP_MAX_CORES  8

typedef struct {
    snd_pcm_extplug_t ext;
    avdsp_b200_t *gpu;
    int32_t *prog;                    /* the program file, as loaded */
    int progWords, dspFormat;
    int dither, timestat, tagoutput, canonical, device;
    int chIn, chOut;                  /* channel counts of the PCM on either side: highest used io index + 1 */
    int nIn, inIdx[32], nOut, outIdx[32];          /* the executor's compact channel maps (ascending io slots) */
    int dense;                        /* the PCM channels ARE the executor's channels, in order: no repacking */
    int nCores, coreFirstOut[AVDSP_MAX_CORES];     /* tagoutput stamps the first output of every core (:132-137) */
    int32_t *packIn, *packOut; size_t packFrames;
    unsigned long lastPeriod;
    int previousSample;
    double spentUs, samples, samplesMax;
} avdsp_b200_plugin_t;

static inline void *area_addr(const snd_pcm_channel_area_t *area, snd_pcm_uframes_t offset) {
    return (char *)area->addr + (area->first + area->step * offset) / 8;
}

static int32_t widen(const void *src, snd_pcm_format_t fmt, size_t pos) {          /* :109-121 */
    if (fmt == SND_PCM_FORMAT_S16) return (int32_t)((uint32_t)(uint16_t)((const int16_t *)src)[pos] << 16);
    if (fmt == SND_PCM_FORMAT_S24_3LE) { const unsigned char *p = (const unsigned char *)src + 3 * pos; return (int32_t)(((uint32_t)p[0] << 8) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 24)); }
    return ((const int32_t *)src)[pos];
}

static snd_pcm_sframes_t b200_transfer(snd_pcm_extplug_t *ext, const snd_pcm_channel_area_t *dst_areas, snd_pcm_uframes_t dst_offset,
                                       const snd_pcm_channel_area_t *src_areas, snd_pcm_uframes_t src_offset, snd_pcm_uframes_t size) {
    avdsp_b200_plugin_t *p = ext->private_data;
    const void *src = area_addr(src_areas, src_offset);
    int32_t *dst = area_addr(dst_areas, dst_offset);
    struct timespec t0, t1;
    if (!p->gpu || size == 0) return p->gpu ? 0 : -EINVAL;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    if (!p->canonical && size != p->lastPeriod) {         /* the old loop nest is core-major inside THIS call's frames */
        if (avdsp_b200_set_order(p->gpu, (int)size) < 0) return -EINVAL;
        p->lastPeriod = size;
    }
    int rc;
    if (p->dense) {
        const int fmt = ext->format == SND_PCM_FORMAT_S16 ? AVDSP_B200_PCM_S16 : ext->format == SND_PCM_FORMAT_S24_3LE ? AVDSP_B200_PCM_S24_3LE : AVDSP_B200_PCM_S32;
        rc = avdsp_b200_process_pcm(p->gpu, src, fmt, dst, (int)size, AVDSP_B200_HOST);
    } else {
        /* a program that skips io slots: the PCM has chIn / chOut channels, the executor takes the used ones only */
        if (size > p->packFrames) {
            free(p->packIn); free(p->packOut);
            p->packIn = malloc(sizeof(int32_t) * size * (p->nIn ? p->nIn : 1));
            p->packOut = malloc(sizeof(int32_t) * size * (p->nOut ? p->nOut : 1));
            if (!p->packIn || !p->packOut) return -ENOMEM;
            p->packFrames = size;
        }
        for (snd_pcm_uframes_t n = 0; n < size; n++)
            for (int k = 0; k < p->nIn; k++)
                p->packIn[n * p->nIn + k] = widen(src, ext->format, n * p->chIn + (p->inIdx[k] - AVDSP_IN_BASE));
        rc = avdsp_b200_process(p->gpu, p->packIn, p->packOut, (int)size, AVDSP_B200_INTERLEAVED, AVDSP_B200_HOST);
        if (rc >= 0)
            for (snd_pcm_uframes_t n = 0; n < size; n++)
                for (int k = 0; k < p->nOut; k++) dst[n * p->chOut + (p->outIdx[k] - AVDSP_OUT_BASE)] = p->packOut[n * p->nOut + k];
    }
    if (rc < 0) { SNDERR("avdsp_b200: %s", avdsp_b200_last_error()); return -EIO; }
    if (p->tagoutput)                                     /* bit-perfect aid: a counter in the low bits of each core's first output, in the old loop order */
        for (int c = 0; c < p->nCores; c++) {
            if (p->coreFirstOut[c] < 0) continue;
            for (snd_pcm_uframes_t n = 0; n < size; n++) {
                int32_t *s = &dst[n * p->chOut + (p->coreFirstOut[c] - AVDSP_OUT_BASE)];
                const int32_t top = (int32_t)((uint32_t)*s & 0xFFFF0000u);
                *s = top | (p->previousSample & 0x0000FF00);
                p->previousSample = (top >> 8) + 0x0100;
            }
        }
    if (p->samplesMax != 0.0) {                           /* :144-160 */
        clock_gettime(CLOCK_MONOTONIC, &t1);
        p->spentUs += (t1.tv_sec - t0.tv_sec) * 1e6 + (t1.tv_nsec - t0.tv_nsec) * 1e-3;
        p->samples += size;
        if (p->samples > p->samplesMax) {
            const double per = p->spentUs / p->samples, slot = 1e6 / (double)ext->rate;
            printf("AVDSP time spent per samples = %f uSec = %f percents at %d hz\n", per, 100.0 * per / slot, ext->rate);
            p->spentUs = 0; p->samples = 0;
            if (p->timestat == 1) p->samplesMax = 0.0;
        }
    }
    return size;
}

static int b200_close(snd_pcm_extplug_t *ext) {
    avdsp_b200_plugin_t *p = ext->private_data;
    if (p) { if (p->gpu) avdsp_b200_destroy(p->gpu); free(p->prog); free(p->packIn); free(p->packOut); free(p); }
    return 0;
}

/* hw params are known: (re)start at the stream's rate -- dspRuntimeReset(fs, 0, dither), :172-189 */
static int b200_init(snd_pcm_extplug_t *ext) {
    avdsp_b200_plugin_t *p = ext->private_data;
    int rc;
    if (!p->gpu) rc = avdsp_b200_create(&p->gpu, p->prog, p->progWords, (int)ext->rate, p->dspFormat, 1, NULL, p->dither, p->device);
    else rc = avdsp_b200_reset(p->gpu, (int)ext->rate, NULL, p->dither);
    if (rc < 0) { SNDERR("avdsp filter not supported sample freq : %d (%s)", ext->rate, avdsp_b200_last_error()); return -EINVAL; }
    p->lastPeriod = 0; p->previousSample = 0; p->spentUs = p->samples = 0.0;
    p->samplesMax = p->timestat ? (double)ext->rate * p->timestat : 0.0;
    if (p->canonical) avdsp_b200_set_order(p->gpu, 0);
    return 0;
}

static const snd_pcm_extplug_callback_t b200_callback = { .transfer = b200_transfer, .init = b200_init, .close = b200_close };

static int load_program(avdsp_b200_plugin_t *p, const char *path) {
    FILE *f = fopen(path, "rb");
    if (!f) return -ENOENT;
    fseek(f, 0, SEEK_END);
    const long bytes = ftell(f);
    fseek(f, 0, SEEK_SET);
    if (bytes < 48 || bytes > (64L << 20)) { fclose(f); return -EINVAL; }
    p->prog = malloc((size_t)bytes);
    p->progWords = (int)(bytes / 4);
    const size_t got = p->prog ? fread(p->prog, 4, (size_t)p->progWords, f) : 0;
    fclose(f);
    return got == (size_t)p->progWords ? 0 : -EIO;
}

SND_PCM_PLUGIN_DEFINE_FUNC(avdsp)
{
    snd_config_iterator_t i, next;
    snd_config_t *sconf = NULL;
    const char *progName = NULL;
    static const unsigned int formats[] = { SND_PCM_FORMAT_S16, SND_PCM_FORMAT_S32, SND_PCM_FORMAT_S24_3LE };
    avdsp_b200_plugin_t *p = calloc(1, sizeof *p);
    if (!p) return -ENOMEM;
    p->ext.version = SND_PCM_EXTPLUG_VERSION;
    p->ext.name = "Avdsp Plugin (B200 batched executor)";
    p->ext.callback = &b200_callback;
    p->ext.private_data = p;
    p->dither = 31;

    snd_config_for_each(i, next, conf) {
        snd_config_t *n = snd_config_iterator_entry(i);
        const char *id, *sv;
        long v;
        if (snd_config_get_id(n, &id) < 0) continue;
        if (!strcmp(id, "comment") || !strcmp(id, "type") || !strcmp(id, "hint")) continue;
        if (!strcmp(id, "slave")) { sconf = n; continue; }
        if (!strcmp(id, "dspprog") && snd_config_get_string(n, &progName) == 0) continue;
        if (!strcmp(id, "dither") && snd_config_get_integer(n, &v) == 0 && (v == 0 || (v >= 7 && v <= 31))) { p->dither = (int)v; continue; }
        if (!strcmp(id, "timestat") && snd_config_get_integer(n, &v) == 0 && v >= 0 && v <= 60) { p->timestat = (int)v; continue; }
        if (!strcmp(id, "tagoutput") && snd_config_get_integer(n, &v) == 0 && v >= 0 && v <= 31) { p->tagoutput = (int)v; continue; }
        if (!strcmp(id, "device") && snd_config_get_integer(n, &v) == 0 && v >= 0 && v < 64) { p->device = (int)v; continue; }
        if (!strcmp(id, "order") && snd_config_get_string(n, &sv) == 0 && (!strcmp(sv, "plugin") || !strcmp(sv, "canonical"))) { p->canonical = !strcmp(sv, "canonical"); continue; }
        SNDERR("Unknown or invalid field %s", id);
        free(p);
        return -EINVAL;
    }
    if (!sconf) { SNDERR("No slave configuration defined for avdsp pcm"); free(p); return -EINVAL; }
    if (!progName) { SNDERR("No dspprog file defined for avdsp pcm"); free(p); return -EINVAL; }
    int err = snd_pcm_extplug_create(&p->ext, name, root, sconf, stream, mode);
    if (err < 0) { free(p); return err; }

    if (load_program(p, progName) < 0) { SNDERR("FATAL ERROR trying to load opcode."); return -EINVAL; }
    /* header / checksum / cores / opcode range, without touching a GPU yet (dspRuntimeInit with fs = 0, :316):
       validate at the lowest rate the program covers; the stream's rate is checked in b200_init */
    static const int rates[] = { 8000, 16000, 24000, 32000, 44100, 48000, 88200, 96000, 176400, 192000, 352800, 384000, 705600, 768000 };
    const int fmin = p->progWords > 8 ? p->prog[7] : -1;
    p->dspFormat = (p->progWords > 6 && (p->prog[6] & 0xFFFF)) ? 2 : 3;       /* Q4.28 integers, else float-encoded (DSP_FORMAT 3) */
    char trace[256];
    if (fmin < 0 || fmin >= 14 ||
        avdsp_b200_describe(p->prog, p->progWords, rates[fmin], p->dspFormat, p->dither, 1, 148, trace, sizeof trace) < 0) {
        SNDERR("FATAL ERROR: problem with opcode header or compatibility (%s)", avdsp_b200_last_error());
        return -EINVAL;
    }
    /* cores and channels exactly as :326-356 derives them: DSP_CORE bitmaps, low 16 slots, count = highest index + 1 */
    unsigned usedIn = 0, usedOut = 0;
    for (p->nCores = 0; p->nCores < AVDSP_MAX_CORES; p->nCores++) {
        const int32_t *core = dspFindCore(p->prog, p->nCores + 1);
        if (!core) break;
        const unsigned in = (unsigned)core[1] & 0xFFFFu, out = (unsigned)core[2] & 0xFFFFu;
        p->coreFirstOut[p->nCores] = -1;
        for (int ch = 0; ch < AVDSP_IO_SLOTS; ch++) if ((out >> ch) & 1u) { p->coreFirstOut[p->nCores] = ch; break; }
        usedIn |= in; usedOut |= out;
        if (core == p->prog) { p->nCores = 1; break; }     /* program without DSP_CORE: one core (not eight) */
    }
    for (int ch = 0; ch < AVDSP_IO_SLOTS; ch++) {
        if (((usedIn >> ch) & 1u) && ch - AVDSP_IN_BASE + 1 > p->chIn) p->chIn = ch - AVDSP_IN_BASE + 1;
        if (((usedOut >> ch) & 1u) && ch - AVDSP_OUT_BASE + 1 > p->chOut) p->chOut = ch - AVDSP_OUT_BASE + 1;
        if ((usedIn >> ch) & 1u) p->inIdx[p->nIn++] = ch;
        if ((usedOut >> ch) & 1u) p->outIdx[p->nOut++] = ch;
    }
    p->dense = p->nIn == p->chIn && p->nOut == p->chOut && (p->nIn == 0 || p->inIdx[0] == AVDSP_IN_BASE) && (p->nOut == 0 || p->outIdx[0] == AVDSP_OUT_BASE);
    printf("AVDSP nbcores %d, nbchanin %d, nbchanout %d\n", p->nCores, p->chIn, p->chOut);

    snd_pcm_extplug_set_param(&p->ext, SND_PCM_EXTPLUG_HW_CHANNELS, p->chIn);
    snd_pcm_extplug_set_slave_param(&p->ext, SND_PCM_EXTPLUG_HW_CHANNELS, p->chOut);
    snd_pcm_extplug_set_param_list(&p->ext, SND_PCM_EXTPLUG_HW_FORMAT, 3, formats);
    snd_pcm_extplug_set_slave_param(&p->ext, SND_PCM_EXTPLUG_HW_FORMAT, SND_PCM_FORMAT_S32);
    *pcmp = p->ext.pcm;
    return 0;
}

SND_PCM_PLUGIN_SYMBOL(avdsp);
