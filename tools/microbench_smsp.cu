// microbench_smsp.cu -- which warps of a CTA share an SM sub-partition (and its quarter-rate IMAD.WIDE pipe)?
// 16 warps per CTA, one CTA per SM; only the warps in `mask` run a MAC loop.  Four active warps that sit on four
// different sub-partitions finish ~4x sooner than four that share one.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ int opaque(int v) { asm volatile("" : "+r"(v)); return v; }
__global__ void __launch_bounds__(512) k(long long* out, int iters, unsigned mask) {
    const int warp = threadIdx.x >> 5;
    if (!((mask >> warp) & 1)) return;
    long long acc[8]; int x[8];
#pragma unroll
    for (int c = 0; c < 8; c++) { acc[c] = threadIdx.x + c; x[c] = threadIdx.x * 3 + c; }
    const int b = 12345 + blockIdx.x;
    for (int i = 0; i < iters; i++)
#pragma unroll
        for (int q = 0; q < 5; q++)
#pragma unroll
            for (int c = 0; c < 8; c++) acc[c] = acc[c] + (long long)opaque(x[c]) * (long long)opaque(b + q);
    long long s = 0;
#pragma unroll
    for (int c = 0; c < 8; c++) s ^= acc[c];
    if (s == 0x123456789abcdefll) out[0] = s;
}
static float run(unsigned mask) {
    long long* d; cudaMalloc(&d, 8);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<<<148, 512>>>(d, 512, mask);
    float best = 1e9f;
    for (int r = 0; r < 3; r++) { cudaEventRecord(a); k<<<148, 512>>>(d, 8192, mask); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; }
    cudaFree(d); return best;
}
int main() {
    struct { const char* name; unsigned mask; } t[] = {
        {"warps 0,1,2,3", 0x000F}, {"warps 0,4,8,12", 0x1111}, {"warps 0,1,4,5", 0x0033}, {"warps 0,2,4,6", 0x0055},
        {"warps 0,5,10,15", 0x8421}, {"warp 0 only", 0x0001}, {"warps 0,4", 0x0011}, {"warps 0,1", 0x0003}, {"warps 0,2", 0x0005}, {"warps 0,8", 0x0101}, {"all 16", 0xFFFF}};
    for (auto& e : t) printf("{\"active\": \"%s\", \"ms\": %.3f}\n", e.name, run(e.mask));
    return 0;
}
