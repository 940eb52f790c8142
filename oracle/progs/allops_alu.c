/*
 * allops_alu: our own DSP program on the unchanged reference encoder API.  TEST INFRASTRUCTURE.
 * Executes every X/Y register opcode and every immediate-value opcode of the runtime at least once
 * (runtime/dsp_runtime.c:337-461 SWAPXY..SQRTX, :643-723 VALUE..AND_VALUE_INT, :390-405 SHIFT),
 * each on its own signal path so that a wrong opcode shows up on one named output.
 *
 * io convention of the linux host: inputs io[8], io[9]; outputs io[0..7] and io[16..].
 * usage: dspcreate -dspprog allops_alu.so -binfile x.bin -dspformat N -fsmin 48000 -fsmax 48000 [-int]
 *   -int : also emit the paths that only make sense for the integer ALU (negative / small SQRTX operands give a
 *          sign-of-NaN that is architecture specific in the float formats).
 * Divisors are constants: the reference divides by whatever Y holds.
 */
#include <stdlib.h>
#include <string.h>
#include "dsp_encoder.h"

#define IN(x)  (8 + (x))

int dspProg(int argc, char **argv) {
    int isint = 0;
    for (int i = 0; i < argc; i++)
        if (strcmp(argv[i], "-int") == 0) isint = 1;

    dsp_PARAM();
    int val1 = dspValue_Default(0.3);
    int gain1 = dspGain_Default(0.6);

    dsp_CORE();
    /* o0: COPYXY, ADDXY */
    dsp_LOAD_GAIN_Fixed(IN(0), 0.5);
    dsp_COPYXY();
    dsp_LOAD_GAIN_Fixed(IN(1), 0.25);
    dsp_ADDXY();
    dsp_SAT0DB();
    dsp_STORE(0);
    /* o1: SUBXY, NEGX */
    dsp_LOAD_GAIN_Fixed(IN(0), 0.5);
    dsp_LOAD_GAIN_Fixed(IN(1), 0.25);
    dsp_SUBXY();
    dsp_NEGX();
    dsp_SAT0DB();
    dsp_STORE(1);
    /* o2: ADDYX, SUBYX, NEGY, COPYYX */
    dsp_LOAD_GAIN_Fixed(IN(0), 0.5);
    dsp_LOAD_GAIN_Fixed(IN(1), 0.25);
    dsp_ADDYX();
    dsp_ADDYX();
    dsp_SUBYX();
    dsp_NEGY();
    dsp_COPYYX();
    dsp_SAT0DB();
    dsp_STORE(2);
    /* o3: CLRXY, VALUE (immediate and PARAM forms), GAIN by PARAM pointer */
    dsp_LOAD_GAIN_Fixed(IN(0), 0.5);
    dsp_CLRXY();
    dsp_VALUE_Fixed(0.125);
    dsp_VALUE(val1);
    dsp_ADDXY();
    if (isint) dsp_SHIFT(31);
    dsp_LOAD_GAIN(IN(1), gain1);
    dsp_ADDXY();
    dsp_SAT0DB();
    dsp_STORE(3);

    dsp_CORE();
    /* o4: VALUE_INT, MULXY, SHIFT right */
    dsp_LOAD(IN(0));
    dsp_VALUE_FixedInt(3);
    dsp_MULXY();
    dsp_SHIFT(-2);
    dsp_STORE(4);
    /* o5: SWAPXY, DIVXY */
    dsp_LOAD(IN(0));
    dsp_VALUE_FixedInt(7);
    dsp_SWAPXY();
    dsp_DIVXY();
    dsp_STORE(5);
    /* o6: DIVYX */
    dsp_LOAD(IN(1));
    dsp_VALUE_FixedInt(-5);
    dsp_DIVYX();
    dsp_COPYYX();
    dsp_STORE(6);
    /* o7: AVGXY */
    dsp_LOAD(IN(0));
    dsp_LOAD(IN(1));
    dsp_AVGXY();
    dsp_STORE(7);
    /* o16: AVGYX */
    dsp_LOAD(IN(0));
    dsp_LOAD(IN(1));
    dsp_NEGX();
    dsp_AVGYX();
    dsp_SWAPXY();
    dsp_STORE(16);

    dsp_CORE();
    /* o17: MUL_VALUE, DIV_VALUE (Q4.28 / float immediates) */
    dsp_LOAD(IN(0));
    dsp_MUL_Fixed(0.6);
    dsp_DIV_Fixed(0.8);
    dsp_STORE(17);
    /* o18: MUL_VALUE_INT, DIV_VALUE_INT, AND_VALUE_INT */
    dsp_LOAD(IN(1));
    dsp_MUL_FixedInt(5);
    dsp_DIV_FixedInt(9);
    dsp_AND_FixedInt(0xFFFFF000);
    dsp_STORE(18);
    /* o19: SHIFT in every form: +-100 (= the mantissa), plain left / right */
    dsp_LOAD(IN(0));
    dsp_SHIFT(100);
    dsp_SHIFT(-3);
    dsp_SHIFT(2);
    dsp_SHIFT(-100);
    dsp_STORE(19);
    /* o20: SQRTX of a square (64-bit branch of the integer ALU) */
    dsp_LOAD(IN(0));
    dsp_COPYXY();
    dsp_MULXY();
    dsp_SQRTX();
    dsp_STORE(20);
    /* o21: SQRTX of a Q59 accumulator */
    dsp_LOAD_GAIN_Fixed(IN(1), 0.5);
    dsp_COPYXY();
    dsp_MULXY();
    if (isint) dsp_SHIFT(-59);
    dsp_SQRTX();
    dsp_GAIN_Fixed(0.9);
    dsp_SAT0DB();
    dsp_STORE(21);
    if (isint) {
        /* o22: SQRTX on raw samples: 32-bit branch for small positives, negatives give 0 */
        dsp_LOAD(IN(0));
        dsp_SHIFT(-9);
        dsp_SQRTX();
        dsp_STORE(22);
        dsp_LOAD(IN(1));
        dsp_SQRTX();
        dsp_STORE(23);
    }
    return dsp_END_OF_CODE();
}
