/*
 * avdsp_oracle.c -- CPU restatement of the AVDSP runtime.  TEST INFRASTRUCTURE ONLY.
 * See avdsp_oracle.h for scope and parity status.
 *
 * Citations are to /root/reference/module_avdsp/runtime/ ("RT/") unless noted.
 * Structure differs from the reference on purpose: state is per instance, the three
 * arithmetic classes (int64 / float / double ALU) are stamped out from one template
 * (avdsp_oracle_exec.inc), and undefined C behaviour the reference relies on is spelled
 * out explicitly (wrap-around adds, x86 shift-count masking) so results do not depend
 * on the compiler.
 */
#include "avdsp_oracle.h"
#include <stdlib.h>
#include <string.h>
#include <math.h>

/* ---- wire format (RT/dsp_header.h:40-132, 213-228) ------------------------------ */
enum {
    OP_END = 0, OP_HEADER, OP_NOP, OP_CORE, OP_PARAM, OP_PARAM_NUM, OP_SERIAL,
    OP_TPDF_CALC, OP_TPDF, OP_WHITE, OP_CLRXY, OP_SWAPXY, OP_COPYXY, OP_COPYYX,
    OP_ADDXY, OP_ADDYX, OP_SUBXY, OP_SUBYX, OP_MULXY, OP_DIVXY, OP_DIVYX, OP_AVGXY, OP_AVGYX,
    OP_NEGX, OP_NEGY, OP_SQRTX, OP_SHIFT, OP_VALUE, OP_VALUE_INT, OP_MUL_VALUE, OP_MUL_VALUE_INT,
    OP_DIV_VALUE, OP_DIV_VALUE_INT, OP_AND_VALUE_INT,
    OP_LOAD, OP_LOAD_GAIN, OP_LOAD_MUX, OP_STORE, OP_LOAD_STORE, OP_LOAD_MEM, OP_STORE_MEM,
    OP_GAIN, OP_SAT0DB, OP_SAT0DB_TPDF, OP_SAT0DB_GAIN, OP_SAT0DB_TPDF_GAIN,
    OP_DELAY_1, OP_DELAY, OP_DELAY_DP, OP_DATA_TABLE, OP_BIQUADS, OP_FIR,
    OP_RMS, OP_DCBLOCK, OP_DITHER, OP_DITHER_NS2, OP_DISTRIB, OP_DIRAC, OP_SQUAREWAVE, OP_CLIP,
    OP_LOAD_MEM_DATA, OP_SINE, OP_MAX
};
enum { HDR_TOTAL = 1, HDR_DATA = 2, HDR_SUM = 3, HDR_CORES = 4, HDR_VERSION = 5, HDR_FMT = 6,
       HDR_FMIN = 7, HDR_FMAX = 8, HDR_IN = 9, HDR_OUT = 10, HDR_HASH = 11 };
#define MANT 28            /* DSP_MANT, DSP_MANTBQ: RT/dsp_header.h:258-267 */
#define NFREQ 14           /* RT/dsp_header.h:136-145 */
#define MAXCORES 32
#define IOMAX 32

static const int kFreqs[NFREQ] = { 8000, 16000, 24000, 32000, 44100, 48000, 88200, 96000,
                                   176400, 192000, 352800, 384000, 705600, 768000 };

static inline int w_op(int32_t w)   { return (int)(((uint32_t)w) >> 16); }
static inline int w_skip(int32_t w) { return (int)(((uint32_t)w) & 0xFFFF); }

/* dither table entry == tpdf_t (RT/dsp_tpdf.h:15-21) */
typedef struct { int dither; int32_t mask; int64_t mask64; int shift; } avo_tpdf;

struct avo_inst {
    int32_t *buf;              /* [code | data] exactly like the reference's single buffer */
    int totalLength, dataSize, format, ncores;
    int coreStart[MAXCORES];   /* word index of first executable opcode of each core */
    uint32_t coreIn[MAXCORES], coreOut[MAXCORES];
    /* what dspRuntimeReset derives (RT/dsp_runtime.c:127-135) */
    int fsIndex, nFreq, bqSkip, bqOffset; uint32_t delayFactor; int rmsFactor;
    /* what dsp_tpdf.h keeps in globals */
    avo_tpdf tg; uint32_t s[4]; int32_t tpdfValue, tpdfRandom; int defaultDither;
};

/* ---- 32-bit float helpers: bit-level restatement of RT/dsp_ieee754.h ----------------- */
static inline uint32_t f2u(float f)  { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float    u2f(uint32_t u){ float f; memcpy(&f, &u, 4); return f; }
static inline uint64_t d2u(double d) { uint64_t u; memcpy(&u, &d, 8); return u; }
static inline double   u2d(uint64_t u){ double d; memcpy(&d, &u, 8); return d; }

/* dspMulFloatFloat, RT/dsp_ieee754.h:336-375: 24x24 mantissa product, truncated, flush to zero */
static float mulFF(float a, float b) {
    uint32_t ua = f2u(a), ub = f2u(b);
    int ea = (ua >> 23) & 255, eb = (ub >> 23) & 255;
    if (ea == 0 || eb == 0) return 0.0f;
    int e = ea + eb - 127;
    if (e < 1) return 0.0f;
    if ((ua ^ ub) & 0x80000000u) e |= 256;
    uint64_t p = (uint64_t)((ua & 0x7FFFFFu) | 0x800000u) * ((ub & 0x7FFFFFu) | 0x800000u);
    uint32_t hi = (uint32_t)(p >> 22);             /* == ((ma<<5)*(mb<<5))>>32 */
    if (hi & (1u << 25)) { e++; hi >>= 2; } else hi >>= 1;
    return u2f((hi & 0x7FFFFFu) | ((uint32_t)e << 23));
}
/* dspMulFloatDouble, :377-410: exact product of two floats as a double */
static double mulFD(float a, float b) {
    uint32_t ua = f2u(a), ub = f2u(b);
    int ea = (ua >> 23) & 255, eb = (ub >> 23) & 255;
    if (ea == 0 || eb == 0) return 0.0;
    int e = 1023 + ea + eb - 254;
    if (e < 1) return 0.0;
    if ((ua ^ ub) & 0x80000000u) e |= 2048;
    uint64_t p = (uint64_t)((ua & 0x7FFFFFu) | 0x800000u) * ((ub & 0x7FFFFFu) | 0x800000u);
    if (p & 0x800000000000ull) { e++; p <<= 5; } else p <<= 6;
    return u2d((p & ((1ull << 52) - 1)) | ((uint64_t)(int64_t)e << 52));
}
/* dspIntToFloatScaled, :204-250: truncating int->float, scaled by 2^-shift.
 * Quirk kept: at most 7 right shifts, so INT_MIN yields -0.5*2^(31-shift)*... (reference bug). */
static float i2fScaled(int32_t x, int shift) {
    if (x == 0) return 0.0f;
    int e = 0;
    uint32_t acc = (uint32_t)x;
    if (x < 0) { acc = 0u - acc; e = 256; }
    int p = 31 - __builtin_clz(acc);
    if (p > 23) { int r = p - 23; if (r > 7) r = 7; acc >>= r; e += r; }
    else        { acc <<= (23 - p); e -= (23 - p); }
    e += 127 + 23 - shift;
    return u2f((acc & 0x7FFFFFu) | ((uint32_t)e << 23));
}
/* dspIntToDoubleScaled, :252-295 (exact) */
static double i2dScaled(int32_t x, int shift) {
    if (x == 0) return 0.0;
    int e = 0;
    uint32_t acc = (uint32_t)x;
    if (x < 0) { acc = 0u - acc; e = 2048; }
    int p = 31 - __builtin_clz(acc);
    acc <<= (31 - p); e -= (31 - p);
    e += 1054 - shift;
    uint64_t m = ((uint64_t)acc << 21) & ((1ull << 52) - 1);
    return u2d(m | ((uint64_t)(int64_t)e << 52));
}
/* dsps31Float0DB, :60-83.  The reference shifts a 32-bit value by n=127-exp which can exceed 31;
 * compiled for x86 the count is taken modulo 32 -- kept, since the oracle restates the binary. */
static int32_t f2s31(float f) {
    uint32_t u = f2u(f);
    int e = (u >> 23) & 255;
    if (e == 0) return 0;
    uint32_t m = ((u & 0x7FFFFFu) | 0x800000u) << 8;
    int n = 127 - e;
    if (n > 0) m >>= (n & 31); else m = 0x7FFFFFFFu;
    if (u & 0x80000000u) m = 0u - m;
    return (int32_t)m;
}
/* dsps31Double0DB, :85-107 (64-bit shift count modulo 64 as on x86-64) */
static int32_t d2s31(double d) {
    uint64_t u = d2u(d);
    int e = (int)((u >> 52) & 2047);
    if (e == 0) return 0;
    int64_t m = (int64_t)((u & ((1ull << 52) - 1)) | (1ull << 52));
    int n = 1044 - e;
    if (n > 21) m >>= (n & 63); else m = 0x7FFFFFFF;
    if ((int64_t)u < 0) m = -m;
    return (int32_t)m;
}
/* dspSaturateFloat0db :170-184 / dspSaturateDouble0db :187-199 */
static float satF(float f) {
    int e = ((int32_t)f2u(f)) >> 23;
    if (e >= 127) return 1.0f;
    if (e < 0 && e >= -129) return -1.0f;
    return f;
}
static double satD(double d) {
    int e = (int)(((int64_t)d2u(d)) >> 52);
    if (e >= 1023) return 1.0;
    if (e < 0 && e >= -1025) return -1.0;
    return d;
}
/* dspShiftFloat :297-314 / dspShiftDouble :316-334: add to the exponent field, no checks */
static float  shiftF(float f, int s)  { return u2f(f2u(f) + ((uint32_t)s << 23)); }
static double shiftD(double d, int s) { return u2d(d2u(d) + ((uint64_t)(int64_t)s << 52)); }
/* dspTruncateFloat0DB :112-138 / dspTruncateDouble0DB :141-167 */
static float truncF(float f, int bit) {
    int32_t i = (int32_t)f2u(f);
    int e = (i >> 23) & 255;
    if (e == 0) return 0.0f;
    int n = 151 - bit - e;
    if (n > 0) {
        if (n >= 24) i = (i >= 0) ? 0 : (int32_t)((uint32_t)(256 + 128 - bit) << 23);
        else { int32_t mask = (int32_t)(0xFFFFFFFFu << n); if (i < 0) i += ~mask; i &= mask; }
    }
    return u2f((uint32_t)i);
}
static double truncD(double d, int bit) {
    int64_t i = (int64_t)d2u(d);
    int e = (int)((i >> 52) & 2047);
    if (e == 0) return 0.0;
    int n = 1076 - bit - e;
    if (n > 0) {
        if (n >= 53) i = (i >= 0) ? 0 : (int64_t)((uint64_t)(uint32_t)((2048 + 1024 - bit) << 20) << 32);
        else { int64_t mask = (int64_t)(~0ull << n); if (i < 0) i += ~mask; i &= mask; }
    }
    return u2d((uint64_t)i);
}

/* ---- fixed-point helpers (RT/dsp_fpmath.h) ------------------------------------------------ */
static inline int64_t wadd(int64_t a, int64_t b) { return (int64_t)((uint64_t)a + (uint64_t)b); }
static inline int64_t wsub(int64_t a, int64_t b) { return (int64_t)((uint64_t)a - (uint64_t)b); }
static inline int64_t wmul(int64_t a, int64_t b) { return (int64_t)((uint64_t)a * (uint64_t)b); }
static inline int64_t mul32(int32_t a, int32_t b) { return (int64_t)a * (int64_t)b; }
static inline int64_t shl64(int64_t a, int n) { return (int64_t)((uint64_t)a << (n & 63)); }
static inline int64_t sar64(int64_t a, int n) { return a >> (n & 63); }
/* dspSaturate64_031, RT/dsp_fpmath.h:84-98 */
static inline int64_t sat64_031(int64_t a) {
    const int64_t lim = (int64_t)1 << (MANT + 31);
    if (a >= lim) return 0x7FFFFFFFll;
    if (a < -lim) return (int64_t)0xFFFFFFFF80000000ull;
    return a >> MANT;
}

/* ---- dither / PRNG (RT/dsp_tpdf.h) ---------------------------------------------------- */
static inline uint32_t rotl32(uint32_t x, unsigned k) { return (x << k) | (x >> (32 - k)); }
/* xoshiro128+, :35-49 */
static uint32_t prng_next(uint32_t *s) {
    uint32_t r = s[0] + s[3], t = s[1] << 9;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl32(s[3], 11);
    return r;
}
/* dspTpdfPrepare :55-80.  returns 1 when `cur` already has this dither, else fills `dst` and returns 0 */
static int tpdf_prepare(const avo_inst *I, const avo_tpdf *cur, avo_tpdf *dst, int dith) {
    if (dith == 0) dith = I->defaultDither;
    if (dith == cur->dither) return 1;
    dst->dither = dith;
    dst->mask   = (int32_t)(0xFFFFFFFFu << ((32 - dith) & 31));   /* -1 << (32-dith), x86 count masking */
    dst->mask64 = (int64_t)((uint64_t)(int64_t)dst->mask << MANT);
    dst->shift  = MANT - dith + 1;
    return 0;
}
/* dspTpdfInit :85-99 */
static void tpdf_init(avo_inst *I, int seed, int defaultDither) {
    I->tpdfRandom = seed; I->tpdfValue = 0; I->defaultDither = defaultDither;
    I->tg.dither = -1;
    tpdf_prepare(I, &I->tg, &I->tg, 0);
    uint32_t u = (uint32_t)seed;
    I->s[0] = u | 1; I->s[1] = rotl32(u | 8, 7); I->s[2] = rotl32(u | 16, 11); I->s[3] = rotl32(u | 24, 17);
}
/* dspTpdfCalc :103-130 (integer part; the per-format conversion is done by the caller) */
static int32_t tpdf_calc(avo_inst *I) {
    int32_t r1 = (int32_t)prng_next(I->s), r2 = (int32_t)prng_next(I->s);
    I->tpdfRandom = r2;
    I->tpdfValue = (r1 >> 1) + (r2 >> 1);
    return I->tpdfValue;
}

/* ---- three ALU classes stamped from one template ---------------------------------- */
#define CLS_INT 1
#define CLS_F32 2
#define CLS_F64 3

#define ACLS CLS_INT
#define EXEC_NAME exec_core_int
#include "avdsp_oracle_exec.inc"
#undef ACLS
#undef EXEC_NAME

#define ACLS CLS_F32
#define EXEC_NAME exec_core_f32
#include "avdsp_oracle_exec.inc"
#undef ACLS
#undef EXEC_NAME

#define ACLS CLS_F64
#define EXEC_NAME exec_core_f64
#include "avdsp_oracle_exec.inc"
#undef ACLS
#undef EXEC_NAME

/* ---- load / validate: dspRuntimeInit RT/dsp_runtime.c:150-195, dspCalcSumCore RT/dsp_header.h:234-251 */
static int find_freq(int fs) { for (int i = 0; i < NFREQ; i++) if (kFreqs[i] == fs) return i; return NFREQ; }

int avo_reset(avo_inst *I, int fs, int seed, int defaultDither) {
    int fi = find_freq(fs);
    if (fi >= NFREQ) return -1;
    int fmin = I->buf[HDR_FMIN], fmax = I->buf[HDR_FMAX];
    if (fi < fmin || fi > fmax) return -2;
    I->fsIndex = fi - fmin;
    I->nFreq = fmax - fmin + 1;
    I->bqSkip = 2 + 6 * I->nFreq;
    I->bqOffset = 5 + 6 * I->fsIndex;
    I->delayFactor = (uint32_t)(4294.967296 * (double)fs);   /* RT/dsp_runtime.c:81-90 */
    I->rmsFactor = (int)(unsigned)(1000.0 / (double)fs);     /* :92-101 (always 0 for fs>1000) */
    memset(I->buf + I->totalLength, 0, sizeof(int32_t) * (size_t)I->dataSize);
    tpdf_init(I, seed, defaultDither);
    return 0;
}

int avo_create(avo_inst **out, const int32_t *prog, int progWords, int maxWords,
               int format, int fs, int seed, int defaultDither) {
    *out = 0;
    if (format < 2 || format > 6) return -7;
    if (progWords < 12 || w_op(prog[0]) != OP_HEADER) return -1;
    int total = prog[HDR_TOTAL], dsz = prog[HDR_DATA];
    if (total < 12 || dsz < 0 || total > progWords) return -6;
    if (total + dsz > maxWords) return -6;
    /* checksum + core count over opcode words only */
    uint32_t sum = 0; int ncore = 0; int p = 0;
    for (;;) {
        int sk = w_skip(prog[p]);
        if (sk == 0) { if (ncore == 0) ncore = 1; break; }
        if (w_op(prog[p]) == OP_CORE) ncore++;
        sum += (uint32_t)prog[p];
        p += sk;
        if (p >= total) return -4;      /* reference prints "BUGG" and stops; a walk off the end never checks out */
    }
    if (ncore < 1) return -3;
    if (sum != (uint32_t)prog[HDR_SUM]) return -4;
    int maxop = (int)(((uint32_t)prog[HDR_FMT]) >> 16), enc = (int)(((uint32_t)prog[HDR_FMT]) & 0xFFFF);
    if (maxop >= OP_MAX) return -5;
    /* The reference would run dspChangeFormat here (unreliable): we require the
     * program to be encoded for the format it is run in. */
    if ((format == 2) != (enc != 0)) return -7;
    if (format == 2 && enc != MANT) return -7;

    avo_inst *I = (avo_inst *)calloc(1, sizeof *I);
    I->buf = (int32_t *)calloc((size_t)(total + dsz + 2), sizeof(int32_t));
    memcpy(I->buf, prog, sizeof(int32_t) * (size_t)total);
    I->totalLength = total; I->dataSize = dsz; I->format = format;
    /* core discovery == dspFindCore + dspFindCoreBegin (RT/dsp_runtime.c:42-77) */
    p = 0; I->ncores = 0;
    for (;;) {
        int sk = w_skip(I->buf[p]);
        if (sk == 0) break;
        if (w_op(I->buf[p]) == OP_CORE && I->ncores < MAXCORES) {
            int c = I->ncores++;
            I->coreIn[c] = (uint32_t)I->buf[p + 1]; I->coreOut[c] = (uint32_t)I->buf[p + 2];
            int q = p;
            for (;;) {
                int o = w_op(I->buf[q]), s2 = w_skip(I->buf[q]);
                if (s2 == 0) break;
                if (o == OP_CORE || o == OP_NOP || o == OP_PARAM || o == OP_PARAM_NUM) q += s2; else break;
            }
            I->coreStart[c] = q;
        }
        p += sk;
    }
    if (I->ncores == 0) {            /* no DSP_CORE: the whole program is one core (dspFindCore :50-52) */
        I->ncores = 1; I->coreStart[0] = 0;
        I->coreIn[0] = (uint32_t)prog[HDR_IN]; I->coreOut[0] = (uint32_t)prog[HDR_OUT];
    }
    *out = I;
    if (fs) { int r = avo_reset(I, fs, seed, defaultDither); if (r) { avo_destroy(I); *out = 0; return r; } }
    return total;
}

void avo_destroy(avo_inst *I) { if (I) { free(I->buf); free(I); } }
int  avo_num_cores(const avo_inst *I) { return I->ncores; }
int  avo_total_length(const avo_inst *I) { return I->totalLength; }
int  avo_data_size(const avo_inst *I) { return I->dataSize; }
int32_t *avo_code(avo_inst *I) { return I->buf; }
int32_t *avo_data(avo_inst *I) { return I->buf + I->totalLength; }
void avo_get_aux(const avo_inst *I, int32_t a[8]) {
    for (int i = 0; i < 4; i++) a[i] = (int32_t)I->s[i];
    a[4] = I->tpdfValue; a[5] = I->tpdfRandom; a[6] = I->tg.dither; a[7] = I->defaultDither;
}
void avo_set_aux(avo_inst *I, const int32_t a[8]) {
    for (int i = 0; i < 4; i++) I->s[i] = (uint32_t)a[i];
    I->tpdfValue = a[4]; I->tpdfRandom = a[5]; I->defaultDither = a[7];
    I->tg.dither = -1; { int d = a[6]; avo_tpdf t = I->tg; tpdf_prepare(I, &t, &I->tg, d ? d : I->defaultDither); }
}

int avo_run_core(avo_inst *I, int core, int32_t *io) {
    if (core < 1 || core > I->ncores) return -1;
    int start = I->coreStart[core - 1];
    switch (I->format) {
    case 2:  return exec_core_int(I, start, io, 1);
    case 3:  return exec_core_f32(I, start, io, 1);
    case 4:  return exec_core_f64(I, start, io, 1);
    case 5:  return exec_core_f32(I, start, io, 0);
    default: return exec_core_f64(I, start, io, 0);
    }
}

int avo_run_frame(avo_inst *I, int32_t *io) {
    for (int c = 1; c <= I->ncores; c++) avo_run_core(I, c, io);
    return 0;
}

int avo_process(avo_inst *I, const int32_t *in, int32_t *out, int nFrames,
                const int *inIdx, int nIn, const int *outIdx, int nOut) {
    int32_t io[IOMAX];
    for (int n = 0; n < nFrames; n++) {
        memset(io, 0, sizeof io);
        for (int c = 0; c < nIn; c++) io[inIdx[c]] = in[(size_t)n * nIn + c];
        avo_run_frame(I, io);
        for (int c = 0; c < nOut; c++) out[(size_t)n * nOut + c] = io[outIdx[c]];
    }
    return 0;
}

/* linux/avdsp_plugin.c:95-142: inputs are io[8+k] for k<nIn, outputs io[k] for k<nOut; per core only
 * its used inputs are filled and only its used outputs are copied; io[] is otherwise indeterminate in the
 * reference (stack garbage) -- zero here. */
int avo_process_plugin_order(avo_inst *I, const int32_t *in, int32_t *out, int nFrames, int period,
                             int nIn, int nOut) {
    int32_t io[IOMAX];
    for (int base = 0; base < nFrames; base += period) {
        int cnt = nFrames - base < period ? nFrames - base : period;
        for (int c = 0; c < I->ncores; c++)
            for (int n = base; n < base + cnt; n++) {
                memset(io, 0, sizeof io);
                for (int ch = 0; ch < 16; ch++)
                    if ((I->coreIn[c] >> ch) & 1u) { int k = ch - 8; if (k >= 0 && k < nIn) io[ch] = in[(size_t)n * nIn + k]; }
                avo_run_core(I, c + 1, io);
                for (int ch = 0; ch < 16; ch++)
                    if ((I->coreOut[c] >> ch) & 1u) { if (ch < nOut) out[(size_t)n * nOut + ch] = io[ch]; }
            }
    }
    return 0;
}
