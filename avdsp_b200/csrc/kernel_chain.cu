// kernel_chain.cu -- systolic executor for programs made of independent signal paths ("chains"):
//     source (LOAD / LOAD_GAIN / LOAD_MUX) -> biquad cascade -> [GAIN] -> SAT0DB[_TPDF][_GAIN]
//     -> [DELAY] -> STORE
// which is what crossovers, EQs and matrix mixers (configs C2, C3, C5) compile to.
//
// Mapping to the B200 (why it looks like this):
//   * The time recurrence of an IIR section cannot be parallelised without changing its
//     fixed-point rounding, so time stays sequential.  Parallelism comes from streams x paths x
//     SECTIONS: one lane owns K consecutive sections of one cascade and keeps their state
//     (64-bit accumulator + x1 x2 y1 y2, reference layout runtime/dsp_biquadSTD.h:45) and their 5
//     coefficients in registers for the whole launch.  A cascade is a systolic pipeline across
//     adjacent lanes of a warp: at step t the lane at depth d works on frame t-d and hands its
//     output to lane d+1 with one __shfl_up.  That is exact because section k of frame n needs only
//     section k-1 of the same frame (dsp_calc_biquads_int, dsp_biquadSTD.h:37-74).
//     4096 streams x 48 sections = 196k lanes -> 41 warps/SM instead of 7 with lane = path.
//   * Everything around the cascades is element-wise over (path, frame) or a block move (delay
//     rings), so it is done a tile of F frames at a time by all threads of the CTA, through shared
//     memory:  phase 1 sources -> x tile;  phase 2 cascades (x tile -> accumulator tile) while one
//     extra warp advances the per-stream dither PRNG;  phase 3 gain/saturate/dither -> delay ->
//     mask -> coalesced store.
//   * A CTA owns NS consecutive streams for the whole launch (state never leaves the SM between
//     tiles); the grid is sized so that CTAs spread evenly over the 148 SMs.
//
// Bit-exactness: all arithmetic goes through avdsp_dev.cuh (mad.wide.s32 products, wrapping adds,
// arithmetic shifts, the asymmetric high-word saturation), identical to the generic kernel.
#include "avdsp_dev.cuh"
#include "kernels.h"
#include <algorithm>
#include <cstdlib>
#include <vector>

namespace avdsp {

constexpr int kPadF = 4;      // row padding (words) of the 32-bit tiles: keeps (channel, frame) reads conflict-free

__device__ __forceinline__ void ctaSync() { __syncthreads(); }

template <int K>
struct LaneState {
    BqStateI s[K];
    int b0[K], b1[K], b2[K], a1[K], a2[K];
};

// one systolic step: input from the previous lane (or the x tile for a head), K sections, output to the
// next lane (and the accumulator tile for a tail).
template <int K, bool PRED>
__device__ __forceinline__ void chainStep(LaneState<K>& L, int& ypass, const int t, const int depth, const int Fv,
                                          const bool head, const bool tail,
                                          const int* __restrict__ xrow, long long* __restrict__ accrow) {
    int x = __shfl_up_sync(0xffffffffu, ypass, 1);
    const int f = t - depth;
    if (PRED) {
        const bool act = (unsigned)f < (unsigned)Fv;
        if (head && act) x = xrow[t];
        LaneState<K> N = L;                       // work on a copy, commit only when this lane is inside the tile
        long long acc = 0;
#pragma unroll
        for (int k = 0; k < K; k++) { x = biquadStepI(N.s[k], x, L.b0[k], L.b1[k], L.b2[k], L.a1[k], L.a2[k]); acc = N.s[k].acc; }
        if (act) {
#pragma unroll
            for (int k = 0; k < K; k++) L.s[k] = N.s[k];
            ypass = x;
            if (tail) accrow[f] = acc;
        }
    } else {
        if (head) x = xrow[t];
        long long acc = 0;
#pragma unroll
        for (int k = 0; k < K; k++) { x = biquadStepI(L.s[k], x, L.b0[k], L.b1[k], L.b2[k], L.a1[k], L.a2[k]); acc = L.s[k].acc; }
        ypass = x;
        if (tail) accrow[f] = acc;
    }
}

template <int K>
__global__ void __launch_bounds__(1024, 1)
k_chain(const __grid_constant__ ChainPlan P, const ChainArgs A, const ChainGeom G) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int NS = G.streamsPerCta, F = G.tileFrames, FP = F + kPadF;
    const int C = P.h.nChains, slots = NS * C;
    const int W = P.h.stateWords;
    // shared memory carve-up
    long long* acc_s = reinterpret_cast<long long*>(smem_raw);            // [slots][F]   phase 1/2 -> 3a
    int* outv_s = reinterpret_cast<int*>(smem_raw);                        // [slots][FP]  alias of acc_s, phase 3b -> 3c
    int* xin_s  = reinterpret_cast<int*>(smem_raw + (size_t)slots * F * 8);// [slots][FP]  phase 1 -> 2; alias post_s 3a -> 3c
    int* post_s = xin_s;
    int* tpdf_s = xin_s + (size_t)slots * FP;                              // [NS][F]
    int* ridx_s = tpdf_s + (size_t)NS * F;                                 // [slots] ring index

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nWork = G.workThreads, nWorkWarps = nWork >> 5;
    const bool isAux = tid >= nWork;                   // the extra warp: dither PRNG, one lane per stream
    const int s0 = blockIdx.x * NS;                    // first stream of this CTA
    const int nsHere = min(NS, A.nStreams - s0);
    const int T = A.nFrames;
    const int storeMask = ditherMask(P.h.storeDither);

    // ---- prologue: section lanes pull coefficients + state into registers --------------------
    LaneState<K> L;
    int depth = 0, slot = -1; bool head = false, tail = false;
    int* stLane = nullptr; int firstSec = 0;
    if (tid < G.laneThreads) {
        const ChainLane e = A.lanes[tid];
        if (e.slot >= 0 && e.slot / C < nsHere) { slot = e.slot; depth = e.depth; head = e.flags & 1; tail = (e.flags & 2) != 0; firstSec = e.firstSec; }
        else depth = e.depth;
    }
#pragma unroll
    for (int k = 0; k < K; k++) {
        L.s[k].acc = 0; L.s[k].x1 = L.s[k].x2 = L.s[k].y1 = L.s[k].y2 = 0;
        L.b0[k] = L.b1[k] = L.b2[k] = L.a1[k] = L.a2[k] = 0;
    }
    if (slot >= 0) {
        const ChainDesc& d = P.chains[slot % C];
        stLane = A.state + (size_t)(s0 + slot / C) * W;
#pragma unroll
        for (int k = 0; k < K; k++) {
            const int sec = firstSec + k;
            const int* cf = P.pool + d.coefOff + 5 * sec;
            L.b0[k] = cf[0]; L.b1[k] = cf[1]; L.b2[k] = cf[2]; L.a1[k] = cf[3]; L.a2[k] = cf[4];
            const int* q = stLane + P.pool[d.secStateOff + sec];
            L.s[k].acc = (long long)(((unsigned long long)(unsigned)q[1] << 32) | (unsigned)q[0]);
            L.s[k].x1 = q[2]; L.s[k].x2 = q[3]; L.s[k].y1 = q[4]; L.s[k].y2 = q[5];
        }
    }
    const int* xrow = xin_s + (size_t)(slot < 0 ? 0 : slot) * FP;
    long long* accrow = acc_s + (size_t)(slot < 0 ? 0 : slot) * F;
    int ypass = 0;

    // aux warp: PRNG registers of stream s0+lane
    Prng g = {0, 0, 0, 0}; int tpdfValue = 0, tpdfRandom = 0, dith = 0; bool drew = false;
    int* auxp = nullptr;
    if (isAux && lane < nsHere) {
        auxp = A.state + (size_t)(s0 + lane) * W + P.h.auxOff;
        g.s0 = auxp[AUX_S0]; g.s1 = auxp[AUX_S1]; g.s2 = auxp[AUX_S2]; g.s3 = auxp[AUX_S3];
        tpdfValue = auxp[AUX_TPDF_VALUE]; tpdfRandom = auxp[AUX_TPDF_RANDOM]; dith = auxp[AUX_DITHER];
    }
    // ring indices -> shared
    for (int i = tid; i < slots; i += blockDim.x) {
        const ChainDesc& d = P.chains[i % C];
        ridx_s[i] = (d.delayN > 0 && i / C < nsHere) ? A.state[(size_t)(s0 + i / C) * W + d.delayOff] : 0;
    }
    ctaSync();

    const int D = G.maxDepth;
    for (int f0 = 0; f0 < T; f0 += F) {
        const int Fv = min(F, T - f0);
        // ================= phase 1: sources -> x tile (or accumulator tile for paths without biquads)
        if (!isAux) {
            const int nblk = (Fv + 31) >> 5;
            for (int u = warp; u < slots * nblk; u += nWorkWarps) {
                const int sl = u / nblk, blk = u - sl * nblk;
                const int strm = sl / C;
                if (strm >= nsHere) continue;
                const ChainDesc& d = P.chains[sl - strm * C];
                const int f = (blk << 5) + lane;
                if (f >= Fv) continue;
                const int* in = A.in + (size_t)(s0 + strm) * A.inStreamStride + (size_t)(f0 + f) * A.inFrameStride;
                long long X;
                if (d.srcKind == SRC_LOAD_MUX) {
                    X = 0;
                    for (int k = 0; k < d.srcCh; k++) {
                        const int ch = P.pool[d.srcArg + 2 * k], gain = P.pool[d.srcArg + 2 * k + 1];
                        const int smp = ch >= 0 ? __ldg(in + (size_t)ch * A.inChStride) : 0;
                        X = mac32(X, smp, gain);
                    }
                    if (f0 + f == T - 1) {       // the reference leaves the last mux value in the data area (dsp_runtime.c:893-896)
                        int* q = A.state + (size_t)(s0 + strm) * W + d.muxStateOff;
                        q[0] = (int)X; q[1] = (int)(X >> 32);
                    }
                } else {
                    const int smp = d.srcCh >= 0 ? __ldg(in + (size_t)d.srcCh * A.inChStride) : 0;
                    X = (d.srcKind == SRC_LOAD_GAIN) ? mul32(smp, d.srcArg) : (long long)smp;
                }
                if (d.nsec > 0) xin_s[(size_t)sl * FP + f] = (int)(X >> kMantBQ);
                else acc_s[(size_t)sl * F + f] = X;
            }
        }
        ctaSync();
        // ================= phase 2: cascades (systolic) || dither PRNG
        if (tid < G.laneThreads) {
            int t = 0;
            const int tFillEnd = min(D - 1, Fv + D - 1);
            for (; t < tFillEnd; t++) chainStep<K, true>(L, ypass, t, depth, Fv, head, tail, xrow, accrow);
#pragma unroll 4
            for (; t < Fv; t++) chainStep<K, false>(L, ypass, t, depth, Fv, head, tail, xrow, accrow);
            for (; t < Fv + D - 1; t++) chainStep<K, true>(L, ypass, t, depth, Fv, head, tail, xrow, accrow);
        } else if (isAux) {
            if (lane < nsHere) {
                int* row = tpdf_s + lane * F;
                if (P.h.hasTpdfCalc) {
                    for (int f = 0; f < Fv; f++) {
                        if (dith == P.h.tpdfDither) { tpdfValue = tpdfDraw(g, tpdfRandom); drew = true; }
                        else dith = P.h.tpdfDither;          // first frame after a reset: table switch, no draw (dsp_runtime.c:539-544)
                        row[f] = tpdfValue;
                    }
                } else {
                    for (int f = 0; f < Fv; f++) row[f] = tpdfValue;
                }
            }
        }
        ctaSync();
        // ================= phase 3a: [gain] -> saturate (+dither, +gain) -> s.31 tile
        if (!isAux) {
            const int nblk = (Fv + 31) >> 5;
            for (int u = warp; u < slots * nblk; u += nWorkWarps) {
                const int sl = u / nblk, blk = u - sl * nblk;
                const int strm = sl / C;
                if (strm >= nsHere) continue;
                const ChainDesc& d = P.chains[sl - strm * C];
                const int f = (blk << 5) + lane;
                if (f >= Fv) continue;
                long long X = acc_s[(size_t)sl * F + f];
                if (d.hasGain) X = X * (long long)d.gainBits;
                if (d.satKind >= SAT_GAIN) { X >>= kMant; X = X * (long long)d.satGainBits; }
                if (d.satKind & 1) X += tpdfScaledI(tpdf_s[strm * F + f], P.h.tpdfShift);
                post_s[(size_t)sl * FP + f] = (int)sat64_031(X);
            }
        }
        ctaSync();
        // ================= phase 3b: delay lines, block form.  Ring layout/positions are the reference's
        // (dsp_runtime.c:769-794): frame j of this launch touches position pos(j); the sample stored there
        // comes back n frames later.
        if (!isAux) {
            const int nblk = (Fv + 31) >> 5;
            for (int u = warp; u < slots * nblk; u += nWorkWarps) {
                const int sl = u / nblk, blk = u - sl * nblk;
                const int strm = sl / C;
                if (strm >= nsHere) continue;
                const ChainDesc& d = P.chains[sl - strm * C];
                if (d.delayN <= 0) continue;
                const int f = (blk << 5) + lane;
                if (f >= Fv) continue;
                const int n = d.delayN, idx = ridx_s[sl];
                const int* ring = A.state + (size_t)(s0 + strm) * W + d.delayOff + 1;
                // a stale index >= n (delay shortened by reload_params) is used once, then wraps to 0
                const int fs = (idx >= n) ? 1 : 0, base = (idx >= n) ? 0 : idx;
                int y;
                if (f < fs) y = ring[idx];
                else {
                    const int gI = f - fs;
                    y = (gI >= n) ? post_s[(size_t)sl * FP + f - n] : ring[(base + gI) % n];
                }
                outv_s[(size_t)sl * FP + f] = y;
            }
        }
        ctaSync();
        // ================= phase 3c: ring update + masked, coalesced output store
        if (!isAux) {
            const int nblk = (Fv + 31) >> 5;
            for (int u = warp; u < slots * nblk; u += nWorkWarps) {
                const int sl = u / nblk, blk = u - sl * nblk;
                const int strm = sl / C;
                if (strm >= nsHere) continue;
                const ChainDesc& d = P.chains[sl - strm * C];
                if (d.delayN <= 0) continue;
                const int f = (blk << 5) + lane;
                if (f >= Fv) continue;
                const int n = d.delayN, idx = ridx_s[sl];
                int* ring = A.state + (size_t)(s0 + strm) * W + d.delayOff + 1;
                const int fs = (idx >= n) ? 1 : 0, base = (idx >= n) ? 0 : idx;
                if (f < fs) ring[idx] = post_s[(size_t)sl * FP + f];
                else if (f + n >= Fv) ring[(base + f - fs) % n] = post_s[(size_t)sl * FP + f];   // not overwritten later in this tile
            }
            const int nOut = P.h.nOut;
            const int total = nsHere * F * nOut;
            for (int i = tid; i < total; i += nWork) {
                const int ch = i % nOut, r = i / nOut;
                const int f = r % F, strm = r / F;
                if (f >= Fv) continue;
                const int c = P.h.chainOfOut[ch];
                int v = 0;
                if (c >= 0) {
                    const int sl = strm * C + c;
                    v = (P.chains[c].delayN > 0 ? outv_s : post_s)[(size_t)sl * FP + f] & storeMask;
                }
                A.out[(size_t)(s0 + strm) * A.outStreamStride + (size_t)(f0 + f) * A.outFrameStride + (size_t)ch * A.outChStride] = v;
            }
        }
        ctaSync();
        // ring indices advance by Fv (dsp_runtime.c:790-792)
        for (int i = tid; i < slots; i += blockDim.x) {
            const int n = P.chains[i % C].delayN;
            if (n > 0) {
                const int idx = ridx_s[i];
                const int fs = (idx >= n) ? 1 : 0, base = (idx >= n) ? 0 : idx;
                ridx_s[i] = (Fv - fs <= 0) ? idx : (base + Fv - fs) % n;
            }
        }
        // (ridx_s is next read in phase 3b, several barriers away)
    }

    // ---- epilogue: registers -> state blocks ---------------------------------------------------
    if (slot >= 0) {
        const ChainDesc& d = P.chains[slot % C];
#pragma unroll
        for (int k = 0; k < K; k++) {
            int* q = stLane + P.pool[d.secStateOff + firstSec + k];
            q[0] = (int)L.s[k].acc; q[1] = (int)(L.s[k].acc >> 32);
            q[2] = L.s[k].x1; q[3] = L.s[k].x2; q[4] = L.s[k].y1; q[5] = L.s[k].y2;
        }
    }
    if (auxp) {
        auxp[AUX_S0] = g.s0; auxp[AUX_S1] = g.s1; auxp[AUX_S2] = g.s2; auxp[AUX_S3] = g.s3;
        auxp[AUX_TPDF_VALUE] = tpdfValue; auxp[AUX_TPDF_RANDOM] = tpdfRandom; auxp[AUX_DITHER] = dith;
        if (drew) {   // TPDF_CALC leaves its last value (as an ALU word) in the data area (dsp_runtime.c:541-543)
            int* q = A.state + (size_t)(s0 + lane) * W + P.h.tpdfDataOff;
            q[0] = tpdfValue; q[1] = tpdfValue >> 31;
        }
    }
    ctaSync();
    for (int i = tid; i < slots; i += blockDim.x) {
        const ChainDesc& d = P.chains[i % C];
        if (d.delayN > 0 && i / C < nsHere) A.state[(size_t)(s0 + i / C) * W + d.delayOff] = ridx_s[i];
    }
}

// ------------------------------------------------------------------------------------------------
bool chainKernelSupports(const ChainPlan& plan) {
    return plan.h.aluClass == ALU_INT64 && plan.h.nChains > 0 && plan.h.nOut > 0 && plan.h.nRaw == 0 && plan.h.nMemCopy == 0 && plan.h.nDelayFirst == 0;
}

static size_t chainSmem(const ChainPlan& p, int NS, int F) {
    const size_t slots = (size_t)NS * p.h.nChains;
    return slots * F * 8 + slots * (F + kPadF) * 4 + (size_t)NS * F * 4 + slots * 4 + 16;
}

// pack the section lanes of NS streams into warps (first-fit decreasing); returns lane threads
static int packLanes(const ChainPlan& p, int NS, int K, ChainLane* out /*may be null*/, int* maxDepth) {
    struct Item { int slot, lanes; };
    std::vector<Item> items;
    const int C = p.h.nChains;
    int md = 0;
    for (int s = 0; s < NS; s++)
        for (int c = 0; c < C; c++) {
            const int n = (p.chains[c].nsec + K - 1) / K;
            if (n > 0) items.push_back({s * C + c, n});
            md = std::max(md, n);
        }
    if (maxDepth) *maxDepth = md;
    std::stable_sort(items.begin(), items.end(), [](const Item& a, const Item& b) { return a.lanes > b.lanes; });
    std::vector<int> fill;                       // lanes used per warp
    std::vector<std::vector<Item>> warps;
    for (const Item& it : items) {
        if (it.lanes > 32) return -1;
        size_t w = 0;
        for (; w < fill.size(); w++) if (fill[w] + it.lanes <= 32) break;
        if (w == fill.size()) { fill.push_back(0); warps.emplace_back(); }
        fill[w] += it.lanes; warps[w].push_back(it);
    }
    const int threads = (int)fill.size() * 32;
    if (out) {
        for (int i = 0; i < threads; i++) { out[i].slot = -1; out[i].depth = 0; out[i].flags = 0; out[i].firstSec = 0; }
        for (size_t w = 0; w < warps.size(); w++) {
            int l = (int)w * 32;
            for (const Item& it : warps[w])
                for (int dpt = 0; dpt < it.lanes; dpt++, l++) {
                    out[l].slot = it.slot; out[l].depth = dpt; out[l].firstSec = dpt * K;
                    out[l].flags = (dpt == 0 ? 1 : 0) | (dpt == it.lanes - 1 ? 2 : 0);
                }
        }
    }
    return threads;
}

static int envInt(const char* name, int dflt) { const char* v = getenv(name); return (v && *v) ? atoi(v) : dflt; }

bool planChainGeometry(const ChainPlan& plan, int nStreams, int numSMs, ChainGeom* geom, ChainLane* lanesOut) {
    // sections per lane: largest K in {4,2,1} dividing every cascade length (override: AVDSP_B200_K)
    int K = envInt("AVDSP_B200_K", 1);
    if (K != 1 && K != 2 && K != 4) K = 1;
    for (int c = 0; c < plan.h.nChains; c++) if (plan.chains[c].nsec % K) K = 1;
    const int F = std::max(32, envInt("AVDSP_B200_F", 64) / 32 * 32);
    const int forceNS = envInt("AVDSP_B200_NS", 0);
    int bestNS = 0, bestThreads = 0, bestDepth = 0; double bestScore = -1;
    const int maxSmemPerSM = 220 * 1024;
    for (int cps = 1; cps <= 4; cps++) {
        int NS = forceNS > 0 ? forceNS : (nStreams + numSMs * cps - 1) / (numSMs * cps);
        NS = std::max(1, std::min(NS, std::min(nStreams, 32)));
        int md = 0;
        const int lt = packLanes(plan, NS, K, nullptr, &md);
        if (lt < 0) return false;
        const int work = std::max(lt, std::min(256, ((NS * plan.h.nChains * F / 4) + 31) / 32 * 32));
        const int threads = work + 32;
        const size_t smem = chainSmem(plan, NS, F);
        if (threads > 1024 || smem > 200 * 1024) continue;
        const int ctas = (nStreams + NS - 1) / NS;
        const int perSM = std::min({cps, (int)(maxSmemPerSM / smem), 1536 / threads > 0 ? 1536 / threads : 1});
        if (perSM < 1) continue;
        const int waves = (ctas + numSMs * perSM - 1) / (numSMs * perSM);
        // throughput proxy: streams in flight per wave over number of waves, favouring more resident lanes
        const double eff = (double)ctas / ((double)waves * numSMs * perSM);
        const double score = eff * std::min(1.0, (double)(perSM * threads) / 1024.0);
        if (score > bestScore) { bestScore = score; bestNS = NS; bestThreads = work; bestDepth = md; }
        if (forceNS > 0) break;
    }
    if (bestNS == 0) return false;
    int md = 0;
    const int lt = packLanes(plan, bestNS, K, lanesOut, &md);
    geom->streamsPerCta = bestNS; geom->secPerLane = K; geom->laneThreads = lt;
    geom->workThreads = std::max(bestThreads, lt); geom->tileFrames = F; geom->maxDepth = std::max(md, 1);
    geom->smemBytes = chainSmem(plan, bestNS, F);
    (void)bestDepth;
    return true;
}

cudaError_t launchChain(const ChainPlan& plan, const ChainGeom& geom, const ChainArgs& args, cudaStream_t stream) {
    const int blocks = (args.nStreams + geom.streamsPerCta - 1) / geom.streamsPerCta;
    const int threads = geom.workThreads + 32;
    cudaError_t e;
#define LAUNCH(KK) \
    e = cudaFuncSetAttribute(k_chain<KK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)geom.smemBytes); \
    if (e != cudaSuccess) return e; \
    k_chain<KK><<<blocks, threads, geom.smemBytes, stream>>>(plan, args, geom);
    switch (geom.secPerLane) {
    case 4: LAUNCH(4); break;
    case 2: LAUNCH(2); break;
    default: LAUNCH(1); break;
    }
#undef LAUNCH
    return cudaGetLastError();
}

} // namespace avdsp
