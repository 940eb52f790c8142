// wire.h -- AVDSP bytecode wire format as consumed by the B200 executor.
//
// Restated from the reference's format definition (the format is the drop-in contract and is
// NOT changed): opcode numbering /root/reference/module_avdsp/runtime/dsp_header.h:40-132,
// opcode word (opcode<<16 | skip) :197-209, 12-word header :213-228, frequency table :136-145,
// Q4.28 parameter mantissa :258-267.
#pragma once
#include <cstdint>

namespace avdsp {

enum Opcode : int {
    OP_END_OF_CODE = 0, OP_HEADER, OP_NOP, OP_CORE, OP_PARAM, OP_PARAM_NUM, OP_SERIAL,
    OP_TPDF_CALC, OP_TPDF, OP_WHITE, OP_CLRXY, OP_SWAPXY, OP_COPYXY, OP_COPYYX,
    OP_ADDXY, OP_ADDYX, OP_SUBXY, OP_SUBYX, OP_MULXY, OP_DIVXY, OP_DIVYX, OP_AVGXY, OP_AVGYX,
    OP_NEGX, OP_NEGY, OP_SQRTX, OP_SHIFT, OP_VALUE, OP_VALUE_INT, OP_MUL_VALUE, OP_MUL_VALUE_INT,
    OP_DIV_VALUE, OP_DIV_VALUE_INT, OP_AND_VALUE_INT,
    OP_LOAD, OP_LOAD_GAIN, OP_LOAD_MUX, OP_STORE, OP_LOAD_STORE, OP_LOAD_MEM, OP_STORE_MEM,
    OP_GAIN, OP_SAT0DB, OP_SAT0DB_TPDF, OP_SAT0DB_GAIN, OP_SAT0DB_TPDF_GAIN,
    OP_DELAY_1, OP_DELAY, OP_DELAY_DP, OP_DATA_TABLE, OP_BIQUADS, OP_FIR,
    OP_RMS, OP_DCBLOCK, OP_DITHER, OP_DITHER_NS2, OP_DISTRIB, OP_DIRAC, OP_SQUAREWAVE, OP_CLIP,
    OP_LOAD_MEM_DATA, OP_SINE,
    OP_MAX_OPCODE
};

// header word indices (dspHeader_t)
enum { H_HEAD = 0, H_TOTAL = 1, H_DATASIZE = 2, H_CHECKSUM = 3, H_NUMCORES = 4, H_VERSION = 5,
       H_FORMAT = 6, H_FREQMIN = 7, H_FREQMAX = 8, H_USEDIN = 9, H_USEDOUT = 10, H_SERIAL = 11,
       H_WORDS = 12 };

constexpr int kMinEncoderVersion = 0x102;   // encoder/dsp_encoder.c:12; older files have other layouts
constexpr int kMant   = 28;   // DSP_MANT
constexpr int kMantBQ = 28;   // DSP_MANTBQ
constexpr int kNumFreq = 14;
constexpr int kFreqTable[kNumFreq] = { 8000, 16000, 24000, 32000, 44100, 48000, 88200, 96000,
                                       176400, 192000, 352800, 384000, 705600, 768000 };
constexpr int kIoSlots = 32;  // dspcreate.c:18 inputOutputMax; the ALSA host uses the low 16

// DSP_FORMAT values (dsp_header.h:11-16)
enum Format { FMT_INT64 = 2, FMT_FLOAT = 3, FMT_DOUBLE = 4, FMT_FLOAT_FLOAT = 5, FMT_DOUBLE_FLOAT = 6 };

inline int wordOpcode(int32_t w) { return (int)((uint32_t)w >> 16); }
inline int wordSkip(int32_t w)   { return (int)((uint32_t)w & 0xFFFFu); }
inline int freqToIndex(int fs) { for (int i = 0; i < kNumFreq; i++) if (kFreqTable[i] == fs) return i; return kNumFreq; }

} // namespace avdsp
