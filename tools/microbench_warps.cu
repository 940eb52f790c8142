// microbench_warps.cu -- how many warps per SM sub-partition does it take to keep the quarter-rate IMAD.WIDE pipe busy?
// One CTA per SM with W warps per sub-partition (4*W warps); every thread runs CH independent accumulating chains
//   acc[c] = IMAD.WIDE(x[c], b, acc[c])        (MODE 0: MACs only)
//   ... plus one funnel shift + one VIADDMNMX per 5 MACs (MODE 1: the shape of a biquad section step in k_chain3)
// Result: IMAD.WIDE per clock per SM as a function of W -- the per-warp issue interval decides how k_chain3 must cut
// cascades into warps.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ int opaque(int v) { asm volatile("" : "+r"(v)); return v; }
__device__ __forceinline__ long long mac32(long long acc, int a, int b) { return acc + (long long)opaque(a) * (long long)opaque(b); }
__device__ __forceinline__ int lo32(long long v) { int l, h; asm("mov.b64 {%0,%1}, %2;" : "=r"(l), "=r"(h) : "l"(v)); return l; }
__device__ __forceinline__ int hi32(long long v) { int l, h; asm("mov.b64 {%0,%1}, %2;" : "=r"(l), "=r"(h) : "l"(v)); return h; }

template <int CH, int MODE>
__global__ void __launch_bounds__(1024) k(long long* out, int iters, int b0) {
    long long acc[CH];
    int x[CH], b[5];
#pragma unroll
    for (int c = 0; c < CH; c++) { acc[c] = (long long)(threadIdx.x + c) * 0x100000001ll; x[c] = threadIdx.x * 7 + c; }
#pragma unroll
    for (int q = 0; q < 5; q++) b[q] = b0 + q + (int)blockIdx.x;
    unsigned worst = 0;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int q = 0; q < 5; q++)
#pragma unroll
            for (int c = 0; c < CH; c++) acc[c] = mac32(acc[c], x[c], b[q]);
        if (MODE == 1) {
#pragma unroll
            for (int c = 0; c < CH; c++) {
                worst = max(worst, (unsigned)hi32(acc[c]) + 0x7fffffeu);
                x[c] = (int)__funnelshift_r((unsigned)lo32(acc[c]), (unsigned)hi32(acc[c]), 28);
            }
        }
    }
    long long s = worst;
#pragma unroll
    for (int c = 0; c < CH; c++) s ^= acc[c] + x[c];
    if (s == 0x123456789abcdefll) out[0] = s;
}

template <int CH, int MODE> void run(int W, int iters) {
    int dev; cudaGetDevice(&dev); cudaDeviceProp p; cudaGetDeviceProperties(&p, dev);
    long long* d; cudaMalloc(&d, 8);
    const int blocks = p.multiProcessorCount, threads = 128 * W;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<CH, MODE><<<blocks, threads>>>(d, iters / 8 + 1, 3);
    double best = 0;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(a); k<CH, MODE><<<blocks, threads>>>(d, iters, 3 + rep); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        const double rate = (double)blocks * threads * CH * 5.0 * iters / (ms * 1e-3);
        if (rate > best) best = rate;
    }
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev);
    const double perClkSM = best / (p.multiProcessorCount * (double)clk * 1e3);
    printf("{\"chains\": %d, \"mode\": \"%s\", \"warps_per_subpartition\": %d, \"T_mac_per_s\": %.3f, \"mac_per_clk_per_SM\": %.2f, "
           "\"cycles_per_warp_mac_per_subpartition\": %.2f, \"cycles_per_mac_per_warp\": %.2f}\n",
           CH, MODE ? "5 mac + shf + viaddmnmx" : "mac only", W, best / 1e12, perClkSM, 128.0 / perClkSM, 128.0 * W / perClkSM);
    cudaFree(d);
}

int main() {
    for (int W : {1, 2, 3, 4, 6, 8}) run<8, 0>(W, 4096);
    for (int W : {1, 2, 3, 4, 6, 8}) run<4, 0>(W, 4096);
    for (int W : {1, 2, 3, 4, 6, 8}) run<8, 1>(W, 4096);
    for (int W : {1, 2, 3, 4, 6, 8}) run<4, 1>(W, 4096);
    for (int W : {1, 2, 3, 4, 6, 8}) run<3, 1>(W, 4096);
    return 0;
}
