// microbench_f32x2.cu -- how do mul.rz.ftz.f32x2 / add.rn.f32x2 (FMUL2 / FADD2) issue on sm_100a?
//
// The float class of k_chain3 (DSP_FORMAT 3) runs its biquad MACs as f32x2 pairs.  ncu (profiles/r2_chain3_c3f_ncu_summary.txt)
// shows neither the issue slots nor the fma pipes saturated.  Question: can ONE warp keep the pipe busy, or does a packed
// instruction hold the warp's issue for longer than its pipe time?
// (Every product takes a value that changes per step: identical `asm volatile` multiplies are merged by the compiler.)
//   mode 0: scalar   acc[i] = acc[i] + acc[i+1] * c   (FMUL.RZ.FTZ + FADD),   K chains
//   mode 1: packed   the same on f32x2 operands (FMUL2 + FADD2),         K independent chains (2K element chains)
//   mode 2: the cascade step: 5 multiplies (independent) + 5 dependent adds per chain, K chains, packed
//   mode 3: mode 2 scalar
// W warps per sub-partition (block = 128 W threads, one block per SM).  Output: cycles per instruction and sub-partition.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long pk(float lo, float hi) { unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) { unsigned long long p; asm volatile("mul.rz.ftz.f32x2 %0, %1, %2;" : "=l"(p) : "l"(a), "l"(b)); return p; }
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) { unsigned long long p; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(p) : "l"(a), "l"(b)); return p; }
__device__ __forceinline__ float mul1(float a, float b) { float p; asm volatile("mul.rz.ftz.f32 %0, %1, %2;" : "=f"(p) : "f"(a), "f"(b)); return p; }
__device__ __forceinline__ float add1(float a, float b) { float p; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(p) : "f"(a), "f"(b)); return p; }

template <int MODE, int K>
__global__ void __launch_bounds__(512) k_bench(float* out, int iters) {
    const float fb = 0.999f;
    float s1[K], c1[5];
    unsigned long long s2[K], c2[5];
#pragma unroll
    for (int i = 0; i < K; i++) { s1[i] = (float)i; s2[i] = pk((float)i, (float)i + 0.5f); }
#pragma unroll
    for (int i = 0; i < 5; i++) { c1[i] = fb - 0.01f * i; c2[i] = pk(fb - 0.01f * i, fb - 0.02f * i); }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
            if (MODE == 0) {
#pragma unroll
                for (int i = 0; i < K; i++) s1[i] = add1(s1[i], mul1(s1[(i + 1) % K], c1[i % 5]));
            } else if (MODE == 1) {
#pragma unroll
                for (int i = 0; i < K; i++) s2[i] = add2(s2[i], mul2(s2[(i + 1) % K], c2[i % 5]));
            } else if (MODE == 2) {
                // K chains; per chain 5 (mul, add): the first multiply takes the chain's previous value (as in the cascade)
                unsigned long long acc[K];
#pragma unroll
                for (int i = 0; i < K; i++) acc[i] = add2(s2[i], mul2(s2[(i + K - 1) % K], c2[0]));
#pragma unroll
                for (int m = 1; m < 5; m++)
#pragma unroll
                    for (int i = 0; i < K; i++) acc[i] = add2(acc[i], mul2(s2[(i + m) % K], c2[m]));      // history terms: operands known a step ahead
#pragma unroll
                for (int i = 0; i < K; i++) s2[i] = acc[i];
            } else {
                float acc[K];
#pragma unroll
                for (int i = 0; i < K; i++) acc[i] = add1(s1[i], mul1(s1[(i + K - 1) % K], c1[0]));
#pragma unroll
                for (int m = 1; m < 5; m++)
#pragma unroll
                    for (int i = 0; i < K; i++) acc[i] = add1(acc[i], mul1(s1[(i + m) % K], c1[m]));
#pragma unroll
                for (int i = 0; i < K; i++) s1[i] = acc[i];
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < K; i++) { s += s1[i]; s += __uint_as_float((unsigned)(s2[i] >> 32)) + __uint_as_float((unsigned)s2[i]); }
    if (s == 123.456f) out[0] = s;
}

template <int MODE, int K> void run(int W, int iters) {
    int dev; cudaGetDevice(&dev); cudaDeviceProp p; cudaGetDeviceProperties(&p, dev);
    float* d; cudaMalloc(&d, 8);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k_bench<MODE, K><<<p.multiProcessorCount, 128 * W>>>(d, iters / 8 + 1);
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(a); k_bench<MODE, K><<<p.multiProcessorCount, 128 * W>>>(d, iters); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
    }
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev);
    const int perUnroll = (MODE < 2 ? 2 * K : 10 * K);                   // pipe instructions per warp and unrolled step
    const double cyc = best * 1e-3 * clk * 1e3 / ((double)iters * 4);   // cycles per unrolled step
    const char* names[] = {"scalar independent", "packed independent", "packed cascade step", "scalar cascade step"};
    printf("{\"mode\": \"%s\", \"chains\": %d, \"warps_per_subpartition\": %d, \"cycles_per_step\": %.2f, \"cycles_per_instr_and_subpartition\": %.3f, "
           "\"element_ops_per_clk_sm\": %.1f}\n", names[MODE], K, W, cyc, cyc / (perUnroll * W),
           (double)perUnroll * W * 4 * 32 * (MODE == 1 || MODE == 2 ? 2 : 1) / cyc);
    cudaFree(d);
}

int main() {
    const int iters = 20000;
    for (int W = 1; W <= 4; W++) {
        run<0, 8>(W, iters); run<1, 4>(W, iters); run<1, 8>(W, iters);
        run<2, 2>(W, iters); run<2, 4>(W, iters); run<2, 8>(W, iters);
        run<3, 4>(W, iters); run<3, 8>(W, iters);
    }
    return 0;
}
