// kernel_mix.cu -- time-parallel executor for programs WITHOUT recurrences: every signal path is
//     source (LOAD / LOAD_GAIN / LOAD_MUX) -> [GAIN] -> SAT0DB[_TPDF][_GAIN] -> [DELAY] -> STORE
// (matrix mixers, routers, delay/gain/dither programs: config C5).  Fixed point (DSP_FORMAT 2), bit-exact.
//
// With no biquad in the path the only things that tie frame n to frame n-1 are (a) the delay lines -- a pure
// shift: the output of frame f is the saturated value of frame f-n -- and (b) the dither PRNG (xoshiro128+,
// runtime/dsp_tpdf.h:35-49), which is LINEAR over GF(2) and can therefore be jumped ahead.  So time is not a
// loop here: the launch is cut into (stream, tile of FT frames) CTAs that stream PCM through shared memory at
// HBM speed; this is the "pure HBM-bound path" (64 B of PCM per frame for the 8x8 mixer).
//
//   k_mix_prng : per stream, J threads each jump to the start of their time segment with a precomputed
//                2L-step transition matrix (128x128 bits) and generate the TPDF values of the segment
//                (dspTpdfCalc, runtime/dsp_tpdf.h:103-130) into a scratch buffer; the last one leaves the PRNG /
//                TPDF state where the reference would (per-stream aux words + the TPDF_CALC data word).
//   k_mix_main : lane = output frame.  A CTA stages the input window [f0-maxDelay, f0+FT) (interleaved PCM is
//                one contiguous run) and the matching dither values, then every output channel evaluates ITS
//                chain at frame f-delay: dense gain-matrix row x 8 inputs (16-byte shared loads, accumulating
//                IMAD.WIDE), gain / dither / saturate, mask, and the frame's outputs leave as 16-byte stores.
//                Each saturated value is computed exactly once (by the output frame that needs it).
//                Frames f < delay read the reference's ring from the state block (previous launch).
//   k_mix_tail : writes the state the reference would hold after T frames: the last `delay` saturated values of
//                every path in ring layout + ring index (dsp_runtime.c:769-794), the last LOAD_MUX value (:893-896).
#include "avdsp_dev.cuh"
#include "kernels.h"
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace avdsp {

__device__ __forceinline__ unsigned mixSmem(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

// ---------------------------------------------------------------------------------------------------------
// saturated s.31 value ("post") of output channel `ch`'s chain at frame fd, from PCM given as a pointer to
// channel 0 of that frame (global or shared), and the dither value of that frame
__device__ __forceinline__ long long mixSourceDense(const MixPlan& M, int ch, const int* __restrict__ fr) {
    // dense gain-matrix row (LOAD / LOAD_GAIN are rows with a single non-zero, LOAD is handled by the caller)
    const int* g = M.mat + ch * kFastTab;
    long long X = 0;
#pragma unroll 4
    for (int k = 0; k < M.nInPad; k++) X = mac32(X, fr[k], g[k]);
    return X;
}
// the same with everything static: `ch` and NP are compile-time, so the gains are constant-bank operands of the
// IMAD.WIDEs and the frame's inputs arrive as 16-byte shared loads
template <int NP>
__device__ __forceinline__ long long mixSourceDenseT(const MixPlan& M, const int ch, const int* __restrict__ fr) {
    long long X = 0;
#pragma unroll
    for (int k4 = 0; k4 < NP; k4 += 4) {
        const int4 v = *reinterpret_cast<const int4*>(fr + k4);
        X = mac32(X, v.x, M.mat[ch * kFastTab + k4 + 0]);
        X = mac32(X, v.y, M.mat[ch * kFastTab + k4 + 1]);
        X = mac32(X, v.z, M.mat[ch * kFastTab + k4 + 2]);
        X = mac32(X, v.w, M.mat[ch * kFastTab + k4 + 3]);
    }
    return X;
}
__device__ __forceinline__ int mixFinish(const MixPlan& M, int flags, int gainBits, int satGainBits, long long X, int tv) {
    if (flags & PF_GAIN) X = X * (long long)gainBits;
    if (flags & PF_SAT_GAIN) { X >>= kMant; X = X * (long long)satGainBits; }
    if (flags & PF_SAT_TPDF) X += tpdfScaledI(tv, M.tpdfShift);
    return sat64_031_s32(X);
}

// ---------------------------------------------------------------------------------------------------------
// segment j of J of stream s: jump to the segment's first draw, generate its dither values, leave the state (last segment)
__device__ __forceinline__ void mixPrngSegment(const MixPlan& M, int* __restrict__ state, int* __restrict__ tpdfBuf,
                                               const unsigned* __restrict__ jump, int s, int j, int T, int J, int L) {
    int* aux = state + (size_t)s * M.stateWords + M.auxOff;
    unsigned st[4] = {(unsigned)aux[AUX_S0], (unsigned)aux[AUX_S1], (unsigned)aux[AUX_S2], (unsigned)aux[AUX_S3]};
    int tpdfValue = aux[AUX_TPDF_VALUE], tpdfRandom = aux[AUX_TPDF_RANDOM];
    const int dith = aux[AUX_DITHER];
    int* row = tpdfBuf + (size_t)s * T;
    if (!M.hasCalc) {                        // no TPDF_CALC in the program: the value never changes
        for (int f = j * L; f < ((j == J - 1) ? T : min(T, (j + 1) * L)); f++) row[f] = tpdfValue;   // the last segment runs to the end
        return;
    }
    // first frame after a reset with another dither width: table switch, no draw, X=0, nothing stored (dsp_runtime.c:539-544)
    const int q = (dith != M.tpdfDither) ? 1 : 0;
    // segment j covers draws [j*L, (j+1)*L) = frames [j*L+q, (j+1)*L+q); the last segment runs to the end
    // jump to the segment start: state <- Mseg^j * state, with Mseg^(2^b) precomputed per level b (columns of the matrix,
    // XOR of those selected by the state's bits; branch-free so that the lanes of a warp stay together)
    for (int lv = 0; (j >> lv) != 0; lv++) {
        if (!((j >> lv) & 1)) continue;
        const uint4* cols = reinterpret_cast<const uint4*>(jump) + lv * 128;
        unsigned n0 = 0, n1 = 0, n2 = 0, n3 = 0;
#pragma unroll 1
        for (int w = 0; w < 4; w++) {
            const unsigned bits = st[w];
#pragma unroll 8
            for (int b = 0; b < 32; b++) {
                const uint4 c = __ldg(cols + w * 32 + b);
                const unsigned m = 0u - ((bits >> b) & 1u);
                n0 ^= c.x & m; n1 ^= c.y & m; n2 ^= c.z & m; n3 ^= c.w & m;
            }
        }
        st[0] = n0; st[1] = n1; st[2] = n2; st[3] = n3;
    }
    Prng g = {st[0], st[1], st[2], st[3]};
    const int d0 = j * L, d1 = (j == J - 1) ? (T - q) : min(T - q, (j + 1) * L);
    if (j == 0 && q && T > 0) row[0] = tpdfValue;
    int d = d0;
    // scalar stores up to a 16-byte boundary, then four values per store (adjacent threads are a whole segment apart)
    for (; d < d1 && (((size_t)(row + d + q)) & 15); d++) { tpdfValue = tpdfDraw(g, tpdfRandom); row[d + q] = tpdfValue; }
    for (; d + 4 <= d1; d += 4) {
        int4 v;
        v.x = tpdfDraw(g, tpdfRandom); v.y = tpdfDraw(g, tpdfRandom); v.z = tpdfDraw(g, tpdfRandom); v.w = tpdfDraw(g, tpdfRandom);
        tpdfValue = v.w;
        *reinterpret_cast<int4*>(row + d + q) = v;
    }
    for (; d < d1; d++) { tpdfValue = tpdfDraw(g, tpdfRandom); row[d + q] = tpdfValue; }
    if (j == J - 1) {
        aux[AUX_S0] = g.s0; aux[AUX_S1] = g.s1; aux[AUX_S2] = g.s2; aux[AUX_S3] = g.s3;
        aux[AUX_TPDF_VALUE] = tpdfValue; aux[AUX_TPDF_RANDOM] = tpdfRandom; aux[AUX_DITHER] = M.tpdfDither;
        if (T - q > 0) {                     // TPDF_CALC leaves its last value (as an ALU word) in the data area (dsp_runtime.c:541-543)
            int* qd = state + (size_t)s * M.stateWords + M.tpdfDataOff;
            qd[0] = tpdfValue; qd[1] = tpdfValue >> 31;
        }
    }
}
__global__ void __launch_bounds__(128)
k_mix_prng(const __grid_constant__ MixPlan M, int* __restrict__ state, int* __restrict__ tpdfBuf, const unsigned* __restrict__ jump,
           int nStreams, int T, int J, int L) {
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    const int s = gid / J, j = gid - s * J;
    if (s >= nStreams) return;
    mixPrngSegment(M, state, tpdfBuf, jump, s, j, T, J, L);
}

// ---------------------------------------------------------------------------------------------------------
constexpr int kMixThreads = 256;

template <int NP>
__global__ void __launch_bounds__(kMixThreads, 4)
k_mix_main(const __grid_constant__ MixPlan M, const MixArgs A) {
    extern __shared__ __align__(16) int smem_mix[];
    const int FT = A.tileFrames, W = M.stateWords, T = A.nFrames;
    constexpr int nInPad = NP;
    const int s = blockIdx.y, f0 = blockIdx.x * FT;
    const int nf = min(FT, T - f0);
    const int w0 = max(0, f0 - M.maxDelay) & ~3;             // first staged frame (multiple of 4: 16-byte aligned rows)
    const int nw = f0 + nf - w0;                             // staged frames
    int* pcm_s = smem_mix;                                   // [nw][nInPad]
    int* tpdf_s = smem_mix + (size_t)A.winFrames * nInPad;   // [nw]
    const int* in = A.in + (size_t)s * A.inStreamStride;
    // ---- stage the window (interleaved PCM: nw*nIn contiguous words)
    if (nInPad == M.nIn && A.vecIn) {
        const int4* src = reinterpret_cast<const int4*>(in + (size_t)w0 * M.nIn);
        int4* dst = reinterpret_cast<int4*>(pcm_s);
        const int n4 = nw * M.nIn / 4;
        for (int i = threadIdx.x; i < n4; i += kMixThreads) dst[i] = __ldg(src + i);
        for (int i = n4 * 4 + threadIdx.x; i < nw * M.nIn; i += kMixThreads) pcm_s[i] = in[(size_t)w0 * M.nIn + i];
    } else {
        for (int i = threadIdx.x; i < nw * nInPad; i += kMixThreads) {
            const int fr = i / nInPad, k = i - fr * nInPad;
            pcm_s[i] = k < M.nIn ? in[(size_t)(w0 + fr) * A.inFrameStride + (size_t)k * A.inChStride] : 0;
        }
    }
    if (M.anyTpdf) {
        const int* tb = A.tpdfBuf + (size_t)s * T + w0;
        for (int i = threadIdx.x; i < nw; i += kMixThreads) tpdf_s[i] = tb[i];
    }
    __syncthreads();

    const int* st = A.state + (size_t)s * W;
    const bool vecOut = A.vecOut != 0;
    if (M.uniform && f0 > M.maxDelay && vecOut) {
        // interior tile of a uniform program (every output: dense row -> SAT0DB_TPDF_GAIN-class finish, same flags):
        // no ring reads, no per-channel decisions; delays are byte offsets from the constant bank
        const bool up = M.tpdfShift >= 0;
        const int sh = (up ? M.tpdfShift : -M.tpdfShift) & 63;
        const int flags = M.uFlags;
        for (int u = threadIdx.x; u < nf; u += kMixThreads) {
            const int f = f0 + u;
            const char* frB = reinterpret_cast<const char*>(pcm_s) + (size_t)(f - w0) * (NP * 4);
            const char* tpB = reinterpret_cast<const char*>(tpdf_s) + (size_t)(f - w0) * 4;
            int* out = A.out + (size_t)s * A.outStreamStride + (size_t)f * A.outFrameStride;
#pragma unroll
            for (int ch0 = 0; ch0 < kFastTab; ch0 += 4) {
                if (ch0 >= M.nOut) break;
                int val[4];
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const int ch = ch0 + q;
                    long long X = mixSourceDenseT<NP>(M, ch, reinterpret_cast<const int*>(frB - M.oDelayPcmBytes[ch]));
                    if (flags & PF_GAIN) X = X * (long long)M.oGain[ch];
                    if (flags & PF_SAT_GAIN) { X >>= kMant; X = X * (long long)M.oSatGain[ch]; }
                    if (flags & PF_SAT_TPDF) {
                        const long long tv = *reinterpret_cast<const int*>(tpB - M.oDelayBytes[ch]);
                        X += up ? (long long)((unsigned long long)tv << sh) : (tv >> sh);
                    }
                    val[q] = sat64_031_s32(X) & M.storeMask;
                }
                *reinterpret_cast<int4*>(out + ch0) = make_int4(val[0], val[1], val[2], val[3]);
            }
        }
        return;
    }
    for (int u = threadIdx.x; u < nf; u += kMixThreads) {
        const int f = f0 + u;
        int* out = A.out + (size_t)s * A.outStreamStride + (size_t)f * A.outFrameStride;
#pragma unroll
        for (int ch0 = 0; ch0 < kFastTab; ch0 += 4) {
            if (ch0 >= M.nOut) break;
            int val[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int ch = ch0 + q;
                int v = 0;                               // outputs no path writes read as 0
                if (ch < M.nOut && M.oChain[ch] >= 0) {
                    const int n = M.oDelay[ch];
                    const int fd = f - n;
                    bool fromRing = fd < 0;
                    int ridx = 0;
                    if (n > 0 && f <= n) {               // frames served by the reference's ring (previous launch), dsp_runtime.c:769-794
                        const int idx0 = st[M.oDelayOff[ch]];
                        if (idx0 >= n || idx0 < 0) {     // stale index: used once, then the ring restarts at 0
                            fromRing = true; ridx = (f == 0) ? idx0 : f - 1;
                        } else if (fd < 0) ridx = (idx0 + f) % n;
                    }
                    if (fromRing) v = st[M.oDelayOff[ch] + 1 + ridx];
                    else {
                        const int* fr = pcm_s + (size_t)(fd - w0) * nInPad;
                        const int flags = M.oFlags[ch];
                        long long X;
                        if (M.oKind[ch] == SRC_RAW) {        // DSP_LOAD_STORE: the sample itself, no saturation, no STORE mask
                            val[q] = M.oSrcCh[ch] >= 0 ? fr[M.oSrcCh[ch]] : 0;
                            continue;
                        }
                        if (M.oKind[ch] == SRC_LOAD) X = M.oSrcCh[ch] >= 0 ? (long long)fr[M.oSrcCh[ch]] : 0ll;
                        else X = mixSourceDenseT<NP>(M, ch, fr);
                        v = mixFinish(M, flags, M.oGain[ch], M.oSatGain[ch], X, (flags & PF_SAT_TPDF) ? tpdf_s[fd - w0] : 0);
                    }
                    v &= M.storeMask;
                }
                val[q] = v;
            }
            if (vecOut) *reinterpret_cast<int4*>(out + ch0) = make_int4(val[0], val[1], val[2], val[3]);
            else {
#pragma unroll
                for (int q = 0; q < 4; q++) if (ch0 + q < M.nOut) out[(size_t)(ch0 + q) * A.outChStride] = val[q];
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// k_mix_stream: the same arithmetic organised per STREAM instead of per (stream, tile): a CTA walks its stream's
// frames in time order, 256 frames per step.  Phase A (lane = frame g): read the frame's PCM once (16-byte loads,
// coalesced), evaluate EVERY output channel's chain on it -- dense gain row (constant-bank operands), gain, dither,
// saturate -- and park the saturated values in a shared-memory ring [channel][frame mod R].  Phase B (lane = output
// frame f): channel ch reads its ring row at f - delay_ch (conflict-free: consecutive lanes, consecutive words), masks,
// and the frame leaves as 16-byte stores.  Every input frame is read from HBM once and mixed once; the ring IS the
// set of delay lines (it is primed from the reference's rings in the state block, dsp_runtime.c:769-794, including
// the stale-index case), so delays cost one shared load per output sample.  One barrier per step: the ring holds
// maxDelay + 2 steps of frames, so phase A of the next step never overwrites what phase B still reads.
// Streams can be cut into time segments when there are too few of them; a segment recomputes maxDelay frames of warm-up.
constexpr int kStreamThreads = 256;

template <int NP, int FLAGS>          // FLAGS >= 0: the program's uniform finish flags at compile time (common programs), -1: read at run time
__global__ void __launch_bounds__(kStreamThreads, 3)
k_mix_stream(const __grid_constant__ MixPlan M, const MixArgs A, const int nSeg, const int segFrames, const int ringMask,
             const unsigned* __restrict__ jump, const int prngL /* > 0: this CTA generates its stream's dither first */, const int fusedTail) {
    extern __shared__ __align__(16) int post_s[];            // [nOut][R]
    constexpr int FT = kStreamThreads;
    const int R = ringMask + 1;
    const int tid = (int)threadIdx.x;
    const int s = (int)blockIdx.x / nSeg, seg = (int)blockIdx.x - s * nSeg;
    const int T = A.nFrames;
    const int fa = seg * segFrames, fb = min(T, fa + segFrames);
    if (fa >= fb) return;
    const int* st = A.state + (size_t)s * M.stateWords;
    const int4* in4 = reinterpret_cast<const int4*>(A.in + (size_t)s * A.inStreamStride);
    const int* tb = M.anyTpdf ? A.tpdfBuf + (size_t)s * T : nullptr;
    int* out = A.out + (size_t)s * A.outStreamStride;
    const int warm = (M.maxDelay + FT - 1) / FT * FT;
    const int g0 = max(0, fa - warm);
    if (prngL > 0) {
        // one stream per CTA (nSeg == 1): the CTA's 256 threads generate the stream's dither values (256 jump-ahead
        // segments) into the scratch row right before using them, so the row is still in L2 when phase A reads it
        mixPrngSegment(M, A.state, A.tpdfBuf, jump, s, tid, T, FT, prngL);
        __syncthreads();
    }
    unsigned staleMask = 0;
    if (g0 == 0) {
        // prime the ring with what the reference's delay rings hold: frame f < n outputs ring[(idx0 + f) % n]; a stale
        // index (>= n after the host shortened the delay) is used once, then the ring restarts at 0
#pragma unroll 1
        for (int ch = 0; ch < M.nOut; ch++) {
            const int n = M.oDelay[ch];
            if (n <= 0) continue;
            const int off = M.oDelayOff[ch];
            const int idx0 = st[off];
            const bool stale = idx0 >= n || idx0 < 0;
            if (stale) staleMask |= 1u << ch;
            int* row = post_s + ch * R;
            for (int i = tid; i < n; i += FT) {
                int v;
                if (stale) v = i == 0 ? st[off + 1 + idx0] : st[off + i];
                else { int r = idx0 + i; if (r >= n) r -= n; v = st[off + 1 + r]; }
                row[(i - n) & ringMask] = v;
            }
        }
    }
    const bool up = M.tpdfShift >= 0;
    const int sh = (up ? M.tpdfShift : -M.tpdfShift) & 63;
    const int flags = FLAGS >= 0 ? FLAGS : M.uFlags;
    int4 cur[NP / 4];
    int ctv = 0;
    {
        const int g = g0 + tid;
        if (g < fb) {
#pragma unroll
            for (int q = 0; q < NP / 4; q++) cur[q] = __ldg(in4 + (size_t)g * (NP / 4) + q);
            if (tb) ctv = __ldg(tb + g);
        }
    }
    for (int gt = g0; gt < fb; gt += FT) {
        const int g = gt + tid;
        // ---- phase A: all chains of frame g
        if (g < fb) {
            const int* v = reinterpret_cast<const int*>(cur);
            int* slotp = post_s + (g & ringMask);
            long long dq = 0;                                 // the frame's dither, scaled once for all channels (dspTpdfApply)
            if (flags & PF_SAT_TPDF) { const long long tv = ctv; dq = up ? (long long)((unsigned long long)tv << sh) : (tv >> sh); }
            // four channels at a time: four independent accumulation chains keep the quarter-rate IMAD.WIDE pipe fed
            // (one chain alone waits ~6 cycles between dependent MACs)
#pragma unroll
            for (int ch0 = 0; ch0 < kFastTab; ch0 += 4) {
                if (ch0 >= M.nOut) break;                     // uniform programs have nOut % 4 == 0
                long long X[4] = {0, 0, 0, 0};
#pragma unroll
                for (int k = 0; k < NP; k++)
#pragma unroll
                    for (int q = 0; q < 4; q++) X[q] = mac32(X[q], v[k], M.mat[(ch0 + q) * kFastTab + k]);
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    long long Y = X[q];
                    if (flags & PF_GAIN) Y = Y * (long long)M.oGain[ch0 + q];
                    if (flags & PF_SAT_GAIN) { Y >>= kMant; Y = Y * (long long)M.oSatGain[ch0 + q]; }
                    Y += dq;
                    slotp[(ch0 + q) * R] = sat64_031_s32(Y);
                }
            }
            if (g == 0 && staleMask) {                        // stale ring index: frame n still outputs ring[n-1], not post[0]
                for (int ch = 0; ch < M.nOut; ch++)
                    if ((staleMask >> ch) & 1u) slotp[ch * R] = st[M.oDelayOff[ch] + M.oDelay[ch]];
            }
        }
        // ---- prefetch the next step's frame while this one drains
        int4 nxt[NP / 4];
        int ntv = 0;
        {
            const int gn = g + FT;
            if (gn < fb) {
#pragma unroll
                for (int q = 0; q < NP / 4; q++) nxt[q] = __ldg(in4 + (size_t)gn * (NP / 4) + q);
                if (tb) ntv = __ldg(tb + gn);
            }
        }
        __syncthreads();
        // ---- phase B: output frame f = g gathers every channel at its own delay
        if (g >= fa && g < fb) {
            int* o = out + (size_t)g * A.outFrameStride;
#pragma unroll
            for (int ch0 = 0; ch0 < kFastTab; ch0 += 4) {
                if (ch0 >= M.nOut) break;
                int val[4];
#pragma unroll
                for (int q = 0; q < 4; q++) val[q] = post_s[(ch0 + q) * R + ((g - M.oDelay[ch0 + q]) & ringMask)] & M.storeMask;
                *reinterpret_cast<int4*>(o + ch0) = make_int4(val[0], val[1], val[2], val[3]);
            }
        }
#pragma unroll
        for (int q = 0; q < NP / 4; q++) cur[q] = nxt[q];
        ctv = ntv;
    }
    if (!fusedTail) return;
    // ---- state after T frames, straight from the ring (nSeg == 1): the reference's delay rings hold the last n saturated
    // values of every path at (idx + j) mod n, the ring index has advanced by T (dsp_runtime.c:769-794), and LOAD_MUX
    // leaves its last 64-bit value in the data area (:893-896).  Same arithmetic as k_mix_tail.
    __syncthreads();
    int* stw = A.state + (size_t)s * M.stateWords;
    for (int c = 0; c < M.nChains; c++) {
        const int ch = M.cOut[c];
        const int n = M.cDelay[c];
        if (M.cMuxOff[c] >= 0 && tid == 0 && T > 0) {
            const int4* fr4 = in4 + (size_t)(T - 1) * (NP / 4);
            long long X = 0;
            for (int q = 0; q < NP / 4; q++) {
                const int4 v = fr4[q];
                X = mac32(X, v.x, M.mat[ch * kFastTab + 4 * q + 0]); X = mac32(X, v.y, M.mat[ch * kFastTab + 4 * q + 1]);
                X = mac32(X, v.z, M.mat[ch * kFastTab + 4 * q + 2]); X = mac32(X, v.w, M.mat[ch * kFastTab + 4 * q + 3]);
            }
            stw[M.cMuxOff[c]] = lo32(X); stw[M.cMuxOff[c] + 1] = hi32(X);
        }
        if (n <= 0) continue;
        const int off = M.cDelayOff[c];
        const int idx0 = stw[off];
        const bool stale = idx0 >= n || idx0 < 0;
        const int idxEff = stale ? n - 1 : idx0;
        const int lo = max(stale ? 1 : 0, T - n);
        __syncthreads();                                   // everyone has read idx0 (and the stale slot) before it is rewritten
        const int* row = post_s + ch * R;
        for (int j = lo + tid; j < T; j += FT) {
            int v = row[j & ringMask];
            stw[off + 1 + (int)(((long long)idxEff + j) % n)] = v;
        }
        if (stale && T >= 1 && tid == 0) {                  // frame 0's own value went to ring[idx0] (the ring slot of frame 0 holds the override)
            const int4* fr4 = in4;
            long long X = 0;
            for (int q = 0; q < NP / 4; q++) {
                const int4 v = fr4[q];
                X = mac32(X, v.x, M.mat[ch * kFastTab + 4 * q + 0]); X = mac32(X, v.y, M.mat[ch * kFastTab + 4 * q + 1]);
                X = mac32(X, v.z, M.mat[ch * kFastTab + 4 * q + 2]); X = mac32(X, v.w, M.mat[ch * kFastTab + 4 * q + 3]);
            }
            stw[off + 1 + idx0] = mixFinish(M, M.oFlags[ch], M.oGain[ch], M.oSatGain[ch], X, ((M.oFlags[ch] & PF_SAT_TPDF) && tb) ? tb[0] : 0);
        }
        __syncthreads();
        if (T > 0 && tid == 0) stw[off] = (int)(((long long)idxEff + T) % n);
    }
}

// ---------------------------------------------------------------------------------------------------------
// state after T frames: delay rings + indices, last mux values.  One CTA per stream; chain c's ring is written
// by frames j in [T-n, T) at position (idx_eff + j) mod n.
__global__ void __launch_bounds__(128)
k_mix_tail(const __grid_constant__ MixPlan M, const MixArgs A) {
    const int s = blockIdx.x, T = A.nFrames, W = M.stateWords;
    int* st = A.state + (size_t)s * W;
    const int* in = A.in + (size_t)s * A.inStreamStride;
    const int* tb = A.tpdfBuf ? A.tpdfBuf + (size_t)s * T : nullptr;
    __shared__ int fr_s[128][kFastTab];
    for (int c = 0; c < M.nChains; c++) {
        const int ch = M.cOut[c];                // an output channel fed by chain c (carries its flattened descriptor)
        const int n = M.cDelay[c];
        const bool mux = M.cMuxOff[c] >= 0;
        if (n <= 0 && !mux) continue;
        const int idx0 = n > 0 ? st[M.cDelayOff[c]] : 0;
        const bool stale = n > 0 && (idx0 >= n || idx0 < 0);
        const int idxEff = stale ? n - 1 : idx0;
        const int lo = max(stale ? 1 : 0, T - n);
        __syncthreads();                         // everyone has read idx0 before lane 0 rewrites it below
        // frames lo..T-1 (ring) and, for a stale index, frame 0 (goes to ring[idx0]); frame T-1 for the mux word
        const int first = (n > 0) ? lo : T - 1;
        for (int j0 = first; j0 < T; j0 += 128) {
            const int j = j0 + threadIdx.x;
            if (j < T) {
                int* fr = fr_s[threadIdx.x];
                for (int k = 0; k < M.nInPad; k++) fr[k] = k < M.nIn ? in[(size_t)j * A.inFrameStride + (size_t)k * A.inChStride] : 0;
                long long X;
                if (M.oKind[ch] == SRC_LOAD) X = M.oSrcCh[ch] >= 0 ? (long long)fr[M.oSrcCh[ch]] : 0ll;
                else X = mixSourceDense(M, ch, fr);
                if (mux && j == T - 1) { st[M.cMuxOff[c]] = lo32(X); st[M.cMuxOff[c] + 1] = hi32(X); }
                if (n > 0) {
                    const int flags = M.oFlags[ch];
                    const int v = mixFinish(M, flags, M.oGain[ch], M.oSatGain[ch], X, ((flags & PF_SAT_TPDF) && tb) ? tb[j] : 0);
                    st[M.cDelayOff[c] + 1 + (int)(((long long)idxEff + j) % n)] = v;
                }
            }
        }
        if (stale && T >= 1 && threadIdx.x == 0) {
            int* fr = fr_s[0];
            for (int k = 0; k < M.nInPad; k++) fr[k] = k < M.nIn ? in[(size_t)k * A.inChStride] : 0;
            long long X;
            if (M.oKind[ch] == SRC_LOAD) X = M.oSrcCh[ch] >= 0 ? (long long)fr[M.oSrcCh[ch]] : 0ll;
            else X = mixSourceDense(M, ch, fr);
            const int flags = M.oFlags[ch];
            st[M.cDelayOff[c] + 1 + idx0] = mixFinish(M, flags, M.oGain[ch], M.oSatGain[ch], X, ((flags & PF_SAT_TPDF) && tb) ? tb[0] : 0);
        }
        __syncthreads();
        if (n > 0 && T > 0 && threadIdx.x == 0) st[M.cDelayOff[c]] = (int)(((long long)idxEff + T) % n);
    }
}

// ---------------------------------------------------------------------------------------------------------
// xoshiro128+ transition matrix over GF(2), as 128 columns of 128 bits (uint4 each), raised to the n-th power
static void xoStep(unsigned s[4]) {
    const unsigned t = s[1] << 9;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = (s[3] << 11) | (s[3] >> 21);
}
struct BitMat { unsigned col[128][4]; };
static void matVec(const BitMat& m, const unsigned v[4], unsigned r[4]) {
    r[0] = r[1] = r[2] = r[3] = 0;
    for (int i = 0; i < 128; i++) if ((v[i >> 5] >> (i & 31)) & 1u) for (int w = 0; w < 4; w++) r[w] ^= m.col[i][w];
}
static void matMul(const BitMat& a, const BitMat& b, BitMat& out) { for (int i = 0; i < 128; i++) matVec(a, b.col[i], out.col[i]); }

void mixJumpMatrix(long long steps, unsigned* out /*[128*4]*/) {
    BitMat base, acc, tmp;
    for (int i = 0; i < 128; i++) {
        unsigned s[4] = {0, 0, 0, 0}; s[i >> 5] = 1u << (i & 31);
        xoStep(s); memcpy(base.col[i], s, 16);
        for (int w = 0; w < 4; w++) acc.col[i][w] = (w == (i >> 5)) ? (1u << (i & 31)) : 0u;      // identity
    }
    while (steps > 0) {
        if (steps & 1) { matMul(base, acc, tmp); acc = tmp; }
        matMul(base, base, tmp); base = tmp;
        steps >>= 1;
    }
    memcpy(out, acc.col, sizeof acc.col);
}

bool buildMixPlan(const ChainPlan& P, MixPlan* M, std::string* why) {
    auto no = [&](const char* w) { if (why) *why = w; return false; };
    if (P.h.aluClass != ALU_INT64) return no("fixed point only");
    if (P.h.nChains <= 0 || P.h.nChains > kFastTab || P.h.nOut <= 0 || P.h.nOut > kFastTab || P.h.nIn > kFastTab) return no("more than 16 paths / channels");
    memset(M, 0, sizeof *M);
    M->nIn = P.h.nIn; M->nOut = P.h.nOut; M->nChains = P.h.nChains;
    M->nInPad = std::max(4, (P.h.nIn + 3) & ~3);
    M->stateWords = P.h.stateWords; M->auxOff = P.h.auxOff;
    M->hasCalc = P.h.hasTpdfCalc; M->tpdfDither = P.h.tpdfDither; M->tpdfDataOff = P.h.tpdfDataOff; M->tpdfShift = P.h.tpdfShift;
    M->storeMask = (int)(0xFFFFFFFFu << ((32 - P.h.storeDither) & 31));
    std::vector<std::vector<int>> chainRow(kFastTab, std::vector<int>(kFastTab, 0));
    for (int c = 0; c < P.h.nChains; c++) {
        const ChainDesc& d = P.chains[c];
        if (d.nsec != 0) return no("a path has biquad sections");
        M->cDelay[c] = d.delayN; M->cDelayOff[c] = d.delayOff; M->cMuxOff[c] = d.srcKind == SRC_LOAD_MUX ? d.muxStateOff : -1;
        M->cOut[c] = d.nStores > 0 ? d.storeCh[0] : 0;
        M->maxDelay = std::max(M->maxDelay, d.delayN);
        if (d.satKind & 1) M->anyTpdf = 1;
        // dense gain row
        int* row = chainRow[c].data();
        if (d.srcKind == SRC_LOAD_MUX) {
            for (int k = 0; k < d.srcCh; k++) {
                const int ch = P.pool[d.srcArg + 2 * k], g = P.pool[d.srcArg + 2 * k + 1];
                if (ch < 0) continue;                    // input the host does not feed: reads 0
                if (row[ch] != 0) return no("a LOAD_MUX lists the same input twice");
                row[ch] = g;
            }
        } else if (d.srcKind == SRC_LOAD_GAIN) { if (d.srcCh >= 0) row[d.srcCh] = d.srcArg; }
    }
    for (int ch = 0; ch < kFastTab; ch++) {
        const int c = ch < P.h.nOut ? P.h.chainOfOut[ch] : -1;
        M->oChain[ch] = c;
        if (c < 0) continue;
        const ChainDesc& d = P.chains[c];
        M->oDelay[ch] = d.delayN; M->oDelayOff[ch] = d.delayOff;
        M->oFlags[ch] = (d.hasGain ? PF_GAIN : 0) | ((d.satKind & 1) ? PF_SAT_TPDF : 0) | (d.satKind >= SAT_GAIN ? PF_SAT_GAIN : 0);
        M->oGain[ch] = d.gainBits; M->oSatGain[ch] = d.satGainBits;
        M->oKind[ch] = d.srcKind; M->oSrcCh[ch] = d.srcCh; M->oMatRow[ch] = ch * kFastTab;
        memcpy(M->mat + ch * kFastTab, chainRow[c].data(), sizeof(int) * kFastTab);     // gain rows are stored per OUTPUT channel
    }
    // "uniform" programs take the branch-free interior path: every output channel is written, by a dense-row chain
    // (LOAD_GAIN / LOAD_MUX), all with the same post-processing flags
    M->uniform = (P.h.nOut & 3) == 0;
    M->uFlags = M->oFlags[0];
    for (int ch = 0; ch < P.h.nOut; ch++) {
        if (M->oChain[ch] < 0 || M->oKind[ch] == SRC_LOAD || M->oKind[ch] == SRC_RAW || M->oFlags[ch] != M->uFlags) M->uniform = 0;
        M->oDelayBytes[ch] = M->oDelay[ch] * 4;
        M->oDelayPcmBytes[ch] = M->oDelay[ch] * M->nInPad * 4;
    }
    return true;
}

// per-stream kernel usable?  (uniform program, interleaved + aligned PCM, ring fits); *nSeg = time segments per stream
static bool mixStreamPath(const MixPlan& M, const MixArgs& A, int numSMs, int* nSegOut, int* segFramesOut, int* ringFramesOut) {
    int ringFrames = 1;
    while (ringFrames < M.maxDelay + 2 * kStreamThreads) ringFrames <<= 1;
    if (!(M.uniform && A.vecIn && A.vecOut && M.nIn == M.nInPad && (size_t)M.nOut * ringFrames * 4 <= 200 * 1024) || getenv("AVDSP_B200_MIX_TILED")) return false;
    const int S = A.nStreams, T = A.nFrames;
    int nSeg = 1;                                            // cut streams into time segments only when there are few of them
    const int want = 6 * numSMs;
    const int minSeg = std::max(4 * M.maxDelay, 8 * kStreamThreads);
    if (S < want) nSeg = std::max(1, std::min((want + S - 1) / S, T / minSeg));
    const int segFrames = ((T + nSeg - 1) / nSeg + kStreamThreads - 1) / kStreamThreads * kStreamThreads;
    *nSegOut = (T + segFrames - 1) / segFrames; *segFramesOut = segFrames; *ringFramesOut = ringFrames;
    return true;
}
// dither PRNG segmentation of a launch: J jump-ahead segments of L draws per stream.  J == kStreamThreads means the
// per-stream kernel generates the values itself (one stream per CTA)
void mixPrngSegments(const MixPlan& M, const MixArgs& A, int numSMs, int* J, int* L) {
    int nSeg, segFrames, ringFrames;
    const int T = A.nFrames;
    if (mixStreamPath(M, A, numSMs, &nSeg, &segFrames, &ringFrames) && nSeg == 1 && T >= 8 * kStreamThreads && !getenv("AVDSP_B200_MIX_SPLIT_PRNG")) *J = kStreamThreads;
    else *J = T >= 4096 ? 64 : (T >= 1024 ? 16 : 1);
    *L = *J > 1 ? (T - 1) / *J : T;
}

cudaError_t launchMix(const MixPlan& M, MixArgs A, const unsigned* dJump, int J, int L, int numSMs, cudaStream_t stream, int* launches) {
    const int S = A.nStreams, T = A.nFrames;
    const bool fusedPrng = J == kStreamThreads;
    int nl = 0;
    if (launches) *launches = 0;
    if ((M.anyTpdf || M.hasCalc) && !fusedPrng) {
        nl++;
        const int th = 128, total = S * J;
        k_mix_prng<<<(total + th - 1) / th, th, 0, stream>>>(M, A.state, A.tpdfBuf, dJump, S, T, J, L);
    }
    cudaError_t e = cudaSuccess;
    // uniform programs (every output: dense row + the same finish) on interleaved, 16-byte aligned PCM: per-stream kernel
    int nSeg = 1, segFrames = 0, ringFrames = 0;
    if (mixStreamPath(M, A, numSMs, &nSeg, &segFrames, &ringFrames)) {
        const size_t ringBytes = (size_t)M.nOut * ringFrames * 4;
        const int prngL = (fusedPrng && (M.anyTpdf || M.hasCalc)) ? L : 0;
        const int fusedTail = nSeg == 1 && !getenv("AVDSP_B200_MIX_SPLIT_TAIL");
#define LAUNCH_STREAM2(NPP, FL) do { \
            e = cudaFuncSetAttribute(k_mix_stream<NPP, FL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ringBytes); \
            if (e != cudaSuccess) return e; \
            k_mix_stream<NPP, FL><<<(unsigned)(S * nSeg), kStreamThreads, ringBytes, stream>>>(M, A, nSeg, segFrames, ringFrames - 1, dJump, prngL, fusedTail); } while (0)
#define LAUNCH_STREAM(NPP) do { \
            if (M.uFlags == (PF_SAT_TPDF | PF_SAT_GAIN)) LAUNCH_STREAM2(NPP, PF_SAT_TPDF | PF_SAT_GAIN); \
            else if (M.uFlags == PF_SAT_TPDF) LAUNCH_STREAM2(NPP, PF_SAT_TPDF); \
            else if (M.uFlags == 0) LAUNCH_STREAM2(NPP, 0); \
            else LAUNCH_STREAM2(NPP, -1); } while (0)
        if (M.nInPad == 4) LAUNCH_STREAM(4); else if (M.nInPad == 8) LAUNCH_STREAM(8); else if (M.nInPad == 12) LAUNCH_STREAM(12); else LAUNCH_STREAM(16);
#undef LAUNCH_STREAM2
#undef LAUNCH_STREAM
        e = cudaGetLastError();
        if (launches) *launches = nl + (fusedTail ? 1 : 2);
        if (e != cudaSuccess || fusedTail) return e;
        k_mix_tail<<<S, 128, 0, stream>>>(M, A);
        return cudaGetLastError();
    }
    // tile length: window (FT + maxDelay) x nInPad words + dither values, two CTAs per SM
    // several CTAs per SM so that one tile's load phase overlaps the others' arithmetic (override: AVDSP_B200_MIX_FT)
    int FT = 1024;
    auto smemFor = [&](int ft) { return (size_t)(ft + M.maxDelay + 4) * (M.nInPad + 1) * 4; };
    while (FT > 256 && smemFor(FT) > 72 * 1024) FT >>= 1;
    if (const char* ev = getenv("AVDSP_B200_MIX_FT")) { const int v = atoi(ev); if (v >= 64 && smemFor(v) <= 200 * 1024) FT = v; }
    A.tileFrames = FT; A.winFrames = FT + M.maxDelay + 4;
    const size_t smem = smemFor(FT);
    dim3 grid((T + FT - 1) / FT, S);
#define LAUNCH_MIX(NPP) do { \
        e = cudaFuncSetAttribute(k_mix_main<NPP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        if (e != cudaSuccess) return e; \
        k_mix_main<NPP><<<grid, kMixThreads, smem, stream>>>(M, A); } while (0)
    if (M.nInPad == 4) LAUNCH_MIX(4); else if (M.nInPad == 8) LAUNCH_MIX(8); else if (M.nInPad == 12) LAUNCH_MIX(12); else LAUNCH_MIX(16);
#undef LAUNCH_MIX
    k_mix_tail<<<S, 128, 0, stream>>>(M, A);
    if (launches) *launches = nl + 2;
    return cudaGetLastError();
}

} // namespace avdsp
