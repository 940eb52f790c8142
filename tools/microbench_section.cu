// microbench_section.cu -- the section lane-step of kernel_chain2.cu in isolation: what limits it?
// 148 CTAs x W warps, every lane owns K=2 sections, 48000 steps, x from shared memory, tails write to shared.
// Variants: 0 = as in the kernel (shfl + vote + branch), 1 = no vote (per-lane branch), 2 = no saturation test,
//           3 = no shfl (every lane reads x from smem), 4 = no shfl + no saturation test (pure MAC chains + shift).
#include <cstdio>
#include <cuda_runtime.h>
#include "../avdsp_b200/csrc/avdsp_dev.cuh"
using namespace avdsp;

constexpr unsigned kSatBias = (1u << (kMantBQ - 1)) - 2u, kSatLimit = (1u << kMantBQ) - 3u;
template <int K> struct Lane2 { long long acc[K]; int x1[K], x2[K], y1[K], y2[K], b0[K], b1[K], b2[K], a1[K], a2[K]; };

template <int K, int V>
__device__ __forceinline__ void laneStep(Lane2<K>& L, int xin) {
    int in[K]; in[0] = xin;
#pragma unroll
    for (int j = 1; j < K; j++) in[j] = L.y1[j - 1];
    long long a[K]; unsigned worst = 0;
#pragma unroll
    for (int j = 0; j < K; j++) {
        long long acc = L.acc[j];
        acc = mac32(acc, L.x1[j], L.b1[j]); acc = mac32(acc, L.x2[j], L.b2[j]);
        acc = mac32(acc, L.y1[j], L.a1[j]); acc = mac32(acc, L.y2[j], L.a2[j]);
        acc = mac32(acc, in[j], L.b0[j]);
        a[j] = acc;
        worst = max(worst, (unsigned)hi32(acc) + kSatBias);
    }
    if (V == 0 || V == 3) { if (__any_sync(0xffffffffu, worst > kSatLimit)) {
#pragma unroll
        for (int j = 0; j < K; j++) a[j] = biquadSat(a[j]); } }
    else if (V == 1) { if (worst > kSatLimit) {
#pragma unroll
        for (int j = 0; j < K; j++) a[j] = biquadSat(a[j]); } }
#pragma unroll
    for (int j = 0; j < K; j++) { L.acc[j] = a[j]; L.x2[j] = L.x1[j]; L.x1[j] = in[j]; L.y2[j] = L.y1[j]; L.y1[j] = q59ToS31(a[j]); }
}

template <int K, int V, int UNR>
__global__ void __launch_bounds__(1024, 1) k(const int* coef, int* out, int steps) {
    __shared__ int xs[64];
    __shared__ int ys[1024];
    Lane2<K> L;
    const int tid = threadIdx.x, lane = tid & 31;
    if (tid < 64) xs[tid] = (tid * 2654435761u) >> 4;
#pragma unroll
    for (int k2 = 0; k2 < K; k2++) {
        L.acc[k2] = 0; L.x1[k2] = L.x2[k2] = L.y1[k2] = L.y2[k2] = 0;
        const int* c = coef + 5 * ((tid * K + k2) % 48);
        L.b0[k2] = c[0]; L.b1[k2] = c[1]; L.b2[k2] = c[2]; L.a1[k2] = c[3]; L.a2[k2] = c[4];
    }
    const bool head = (lane % 3) == 0, tail = (lane % 3) == 2;
    __syncthreads();
    for (int t0 = 0; t0 < steps; t0 += 32) {
#pragma unroll 1
        for (int j0 = 0; j0 < 32; j0 += UNR) {
#pragma unroll
            for (int jj = 0; jj < UNR; jj++) {
                const int j = j0 + jj;
                int x;
                if (V < 3) { x = __shfl_up_sync(0xffffffffu, L.y1[K - 1], 1); if (head) x = xs[j]; }
                else x = xs[j] + lane;
                laneStep<K, V>(L, x);
                if (tail) ys[tid] = L.y1[K - 1];
            }
        }
    }
    int s = 0;
#pragma unroll
    for (int k2 = 0; k2 < K; k2++) s += (int)L.acc[k2] + L.y1[k2];
    out[blockIdx.x * blockDim.x + tid] = s + ys[tid];
}

template <int V, int UNR> void run(int warps, const int* dcoef, int* dout) {
    const int steps = 48000, blocks = 148;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<2, V, UNR><<<blocks, warps * 32>>>(dcoef, dout, 3200);
    float best = 1e9;
    for (int r = 0; r < 2; r++) {
        cudaEventRecord(a); k<2, V, UNR><<<blocks, warps * 32>>>(dcoef, dout, steps); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
    }
    const double macs = (double)blocks * warps * 32 * 2 * 5 * steps;
    printf("{\"variant\": %d, \"unroll\": %d, \"warps\": %d, \"ms\": %.3f, \"T_mac_per_s\": %.3f}\n", V, UNR, warps, best, macs / (best * 1e-3) / 1e12);
}

int main() {
    // Butterworth-ish Q4.28 coefficients (stable): b0 b1 b2 a1-1 a2
    int h[48 * 5];
    for (int i = 0; i < 48; i++) { h[5*i] = 2000000 + 1000 * i; h[5*i+1] = 4000000; h[5*i+2] = 2000000; h[5*i+3] = 250000000 - 100000 * i; h[5*i+4] = -240000000 + 90000 * i; }
    int *dc, *dout; cudaMalloc(&dc, sizeof h); cudaMemcpy(dc, h, sizeof h, cudaMemcpyHostToDevice); cudaMalloc(&dout, 148 * 1024 * 4);
    for (int w : {21, 28, 32, 12}) { run<0, 8>(w, dc, dout); }
    run<0, 32>(21, dc, dout);
    run<1, 8>(21, dc, dout);
    run<2, 8>(21, dc, dout);
    run<3, 8>(21, dc, dout);
    run<4, 8>(21, dc, dout);
    return 0;
}
