// microbench_mix.cu -- how many other instructions issue in the shadow of a quarter-rate IMAD.WIDE on sm_100a?
// Per thread: 4 independent accumulate chains (IMAD.WIDE, multiplicand = running low word) and NALU
// independent ALU instructions (LOP3 / IADD3 / SHF flavours) per IMAD.WIDE on separate registers.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ int lo32(long long v) { int l, h; asm("mov.b64 {%0,%1}, %2;" : "=r"(l), "=r"(h) : "l"(v)); (void)h; return l; }

template <int NALU, int KIND>
__global__ void __launch_bounds__(256) k(long long* out, int iters, int b0) {
    long long acc[4]; unsigned r[4][4];
    const int b = b0 + (int)blockIdx.x;
    __shared__ int smbuf[256 * 16];
    const unsigned sm = (unsigned)__cvta_generic_to_shared(smbuf) + threadIdx.x * 64;
#pragma unroll
    for (int c = 0; c < 4; c++) { acc[c] = (long long)(threadIdx.x + c) * 0x100000001ll;
#pragma unroll
        for (int q = 0; q < 4; q++) r[c][q] = threadIdx.x * 7 + c * 13 + q; }
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int c = 0; c < 4; c++) {
            acc[c] = acc[c] + (long long)lo32(acc[c]) * (long long)b;
#pragma unroll
            for (int q = 0; q < NALU; q++) {
                if (KIND == 0) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[c][q]) : "r"(b), "r"(i));          // LOP3
                else if (KIND == 1) asm volatile("add.u32 %0, %0, %1;" : "+r"(r[c][q]) : "r"(b));                        // IADD3 (or IMAD.IADD at ptxas' whim)
                else if (KIND == 2) asm volatile("shf.r.wrap.b32 %0, %0, %1, 7;" : "+r"(r[c][q]) : "r"(b));               // SHF
                else if (KIND == 3) asm volatile("mad.lo.u32 %0, %0, %1, %0;" : "+r"(r[c][q]) : "r"(b));                  // IMAD (fma pipe)
                else if (KIND == 4) asm volatile("{ .reg .pred p; setp.gt.u32 p, %0, %1; @p add.u32 %0, %0, 1; }" : "+r"(r[c][q]) : "r"(b));   // ISETP + predicated add
                else if (KIND == 5) asm volatile("max.u32 %0, %0, %1;" : "+r"(r[c][q]) : "r"(b + i));                    // VIMNMX
                else if (KIND == 6) asm volatile("shfl.sync.up.b32 %0, %0, 1, 0, 0xffffffff;" : "+r"(r[c][q]));           // SHFL
                else if (KIND == 7) asm volatile("ld.shared.b32 %0, [%1];" : "=r"(r[c][q]) : "r"(sm + 4 * q + 16 * c) : "memory"); // LDS
                else if (KIND == 8) asm volatile("st.shared.b32 [%1], %0;" :: "r"(r[c][q]), "r"(sm + 4 * q + 16 * c) : "memory"); // STS
                else if (KIND == 9) asm volatile("mov.b32 %0, %1;" : "=r"(r[c][q]) : "r"(r[c][(q + 1) & 3]));             // MOV (or IMAD.MOV)
                else if (KIND == 10) asm volatile("{ .reg .pred p; setp.gt.u32 p, %0, %1; selp.b32 %0, %0, %1, p; }" : "+r"(r[c][q]) : "r"(b));   // ISETP + SEL
                else if (KIND == 11) asm volatile("{ .reg .pred p, q; setp.gt.u32 p, %0, %1; vote.sync.any.pred q, p, 0xffffffff; @q add.u32 %0, %0, 1; }" : "+r"(r[c][q]) : "r"(b)); // ISETP + VOTE + pred add
                else if (KIND == 12) asm volatile("prmt.b32 %0, %0, %1, 0x3210;" : "+r"(r[c][q]) : "r"(b));               // PRMT

            }
        }
    }
    long long s = 0;
#pragma unroll
    for (int c = 0; c < 4; c++) { s ^= acc[c];
#pragma unroll
        for (int q = 0; q < 4; q++) s += r[c][q]; }
    if (s == 0x123456789abcdefll) out[0] = s;
}

template <int NALU, int KIND> void run(const char* kind) {
    int dev; cudaGetDevice(&dev); cudaDeviceProp p; cudaGetDeviceProperties(&p, dev);
    long long* d; cudaMalloc(&d, 8);
    const int blocks = p.multiProcessorCount * 4, threads = 256, iters = 8192;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<NALU, KIND><<<blocks, threads>>>(d, 100, 3);
    float best = 1e9;
    for (int rep = 0; rep < 2; rep++) { cudaEventRecord(a); k<NALU, KIND><<<blocks, threads>>>(d, iters, 3 + rep); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; }
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev);
    const double wide = (double)blocks * threads * 4.0 * iters;
    const double cyc = best * 1e-3 * clk * 1e3;                                  // cycles elapsed
    const double warpsPerSmsp = (double)blocks * threads / 32 / p.multiProcessorCount / 4;
    printf("{\"other\": \"%s\", \"n_other_per_wide\": %d, \"ms\": %.3f, \"T_wide_per_s\": %.2f, \"cycles_per_wide_per_smsp\": %.2f}\n", kind, NALU, best,
           wide / (best * 1e-3) / 1e12, cyc / (4.0 * iters * warpsPerSmsp));
    cudaFree(d);
}

int main() {
    run<0, 0>("none");
    run<1, 0>("LOP3"); run<2, 0>("LOP3"); run<3, 0>("LOP3"); run<4, 0>("LOP3");
    run<1, 1>("ADD"); run<2, 1>("ADD"); run<3, 1>("ADD"); run<4, 1>("ADD");
    run<1, 2>("SHF"); run<2, 2>("SHF"); run<4, 2>("SHF");
    run<1, 3>("IMAD"); run<2, 3>("IMAD");
    run<1, 4>("ISETP+@ADD"); run<2, 4>("ISETP+@ADD");
    run<1, 5>("VIMNMX"); run<2, 5>("VIMNMX"); run<4, 5>("VIMNMX");
    run<1, 6>("SHFL"); run<2, 6>("SHFL");
    run<1, 7>("LDS"); run<2, 7>("LDS"); run<4, 7>("LDS");
    run<1, 8>("STS"); run<2, 8>("STS"); run<4, 8>("STS");
    run<1, 9>("MOV"); run<2, 9>("MOV"); run<4, 9>("MOV");
    run<1, 10>("ISETP+SEL"); run<2, 10>("ISETP+SEL");
    run<1, 11>("ISETP+VOTE+@ADD"); run<2, 11>("ISETP+VOTE+@ADD");
    run<1, 12>("PRMT"); run<2, 12>("PRMT"); run<4, 12>("PRMT");
    return 0;
}
