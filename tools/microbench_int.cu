// microbench_int.cu -- integer-pipe rates that bound the fixed-point biquad kernels on B200 (sm_100a).
// Each kernel runs 8 independent data-dependent chains per thread; the multiplicand of every MAC is the low
// word of the running accumulator (nothing can be hoisted).  Check the SASS (cuobjdump -sass) before
// trusting a number: the variants differ only in which instruction mix ptxas is steered to.
//   fused : acc = IMAD.WIDE(a, b, acc)                    (signed 32x32+64 multiply-accumulate)
//   split : p = IMAD.WIDE(a, b, 0); acc = IADD3/IADD3.X   (what ptxas emits for mad.wide.s32 chains it re-associates)
//   add64 : acc += c  (IADD3 + IADD3.X)
//   imad  : 32-bit IMAD
//   mix   : 5 fused IMAD.WIDE + 3 ALU ops (shape of one biquad section)
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ int lo32(long long v) { int l, h; asm("mov.b64 {%0,%1}, %2;" : "=r"(l), "=r"(h) : "l"(v)); return l; }
__device__ __forceinline__ int hi32(long long v) { int l, h; asm("mov.b64 {%0,%1}, %2;" : "=r"(l), "=r"(h) : "l"(v)); return h; }

template <int MODE>
__global__ void __launch_bounds__(256) k(long long* out, int iters, int b0) {
    long long acc[8];
    const int b = b0 + (int)blockIdx.x;
    const long long cst = (long long)b * 0x10001ll + threadIdx.x;
#pragma unroll
    for (int k2 = 0; k2 < 8; k2++) acc[k2] = (long long)(threadIdx.x + k2) * 0x100000001ll;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k2 = 0; k2 < 8; k2++) {
            if (MODE == 0) {            // fused mad.wide, C++ form (NVVM -> mad.wide.s32 -> IMAD.WIDE R,R,R,R)
                acc[k2] = acc[k2] + (long long)lo32(acc[k2]) * (long long)b;
            } else if (MODE == 1) {     // mul.wide + 64-bit add kept apart
                long long p; asm volatile("mul.wide.s32 %0, %1, %2;" : "=l"(p) : "r"(lo32(acc[k2])), "r"(b));
                asm volatile("add.s64 %0, %0, %1;" : "+l"(acc[k2]) : "l"(p));
            } else if (MODE == 2) {     // 64-bit add only
                asm volatile("add.s64 %0, %0, %1;" : "+l"(acc[k2]) : "l"(cst));
            } else if (MODE == 3) {     // 32-bit IMAD
                int v = lo32(acc[k2]); asm volatile("mad.lo.s32 %0, %0, %1, %0;" : "+r"(v) : "r"(b)); acc[k2] = v;
            } else if (MODE == 4) {     // one biquad-section-like group: acc + 5 products, saturation test, funnel shift
                long long a = acc[k2];
                const int x = lo32(a), y = hi32(a);
                asm("mad.wide.s32 %0, %1, %2, %0;" : "+l"(a) : "r"(x), "r"(b));
                asm("mad.wide.s32 %0, %1, %2, %0;" : "+l"(a) : "r"(y), "r"(b + 1));
                asm("mad.wide.s32 %0, %1, %2, %0;" : "+l"(a) : "r"(x ^ 5), "r"(b + 2));
                asm("mad.wide.s32 %0, %1, %2, %0;" : "+l"(a) : "r"(y ^ 9), "r"(b + 3));
                asm("mad.wide.s32 %0, %1, %2, %0;" : "+l"(a) : "r"(x ^ y), "r"(b + 4));
                const unsigned t = (unsigned)hi32(a) + 0x7fffffeu;
                if (t > 0xffffffdu) a = 0x07ffffffffffffffll;
                acc[k2] = a;
            } else if (MODE == 5) {     // 5 strictly dependent fused MACs (multiplicand = running low word)
                long long a = acc[k2];
#pragma unroll
                for (int q = 0; q < 5; q++) a = a + (long long)lo32(a) * (long long)b;
                acc[k2] = a;
            }
        }
    }
    long long s = 0;
#pragma unroll
    for (int k2 = 0; k2 < 8; k2++) s ^= acc[k2];
    if (s == 0x123456789abcdefll) out[0] = s;
}

template <int MODE> double run(const char* name, double opsPerInner, int iters) {
    int dev; cudaGetDevice(&dev); cudaDeviceProp p; cudaGetDeviceProperties(&p, dev);
    long long* d; cudaMalloc(&d, 8);
    const int blocks = p.multiProcessorCount * 8, threads = 256;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<MODE><<<blocks, threads>>>(d, iters / 8 + 1, 3);
    double best = 0;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(a); k<MODE><<<blocks, threads>>>(d, iters, 3 + rep); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        const double rate = (double)blocks * threads * 8.0 * iters * opsPerInner / (ms * 1e-3);
        if (rate > best) best = rate;
    }
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev);
    printf("{\"bench\": \"%s\", \"Tops_per_s\": %.3f, \"per_clk_per_SM_at_max_clock\": %.1f}\n", name, best / 1e12,
           best / (p.multiProcessorCount * (double)clk * 1e3));
    cudaFree(d);
    return best;
}

int main() {
    run<0>("imad_wide_fused (mad.wide.s32 acc)", 1, 4096);
    run<1>("imad_wide_mul + add64", 1, 4096);
    run<2>("add64 (IADD3+IADD3.X)", 1, 4096);
    run<3>("imad32", 1, 4096);
    run<4>("section-like: acc + 5 products + sat test, as ptxas schedules it (MACs counted)", 5, 2048);
    run<5>("5 dependent fused MACs per chain step (MACs counted)", 5, 2048);
    return 0;
}
