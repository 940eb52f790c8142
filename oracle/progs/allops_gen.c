/*
 * allops_gen: our own DSP program on the unchanged reference encoder API.  TEST INFRASTRUCTURE.
 * Executes every stateful / generator opcode of the runtime at least once, each on its own output:
 *   TPDF_CALC :537, TPDF :547 (core-local dither tables), WHITE :664, DITHER :1112, DITHER_NS2 :1138,
 *   DCBLOCK :1063, RMS / PWRXY :972 (with and without the moving-average line), DIRAC :1213, SQUAREWAVE :1234,
 *   DATA_TABLE :900, CLIP :1264, DISTRIB :1175, DELAY_1 :726, DELAY_DP :798 (fixed and PARAM forms), SERIAL :863
 *   (all line numbers: runtime/dsp_runtime.c).
 *
 * io convention of the linux host: inputs io[8], io[9]; outputs io[0..7] and io[16..].
 * usage: dspcreate -dspprog allops_gen.so -binfile x.bin -dspformat N -fsmin 44100 -fsmax 192000
 * (DITHER_NS2 insists on a 44.1k..192k table, encoder/dsp_encoder.c:1479).
 */
#include <stdlib.h>
#include <string.h>
#include "dsp_encoder.h"

#define IN(x)  (8 + (x))

static float noiseshaper[] = {          /* triplets per fs, 44.1k .. 192k (the table of dspprogs/testfunction.c:27-33) */
    2.51758, -2.01206, 0.57800,
    2.56669, -2.04479, 0.57800,
    2.75651, -2.50072, 0.77760,
    2.76821, -2.51152, 0.77760,
    2.78567, -2.58690, 0.80595,
    2.78695, -2.59168, 0.80757 };

int dspProg(int argc, char **argv) {
    (void)argc; (void)argv;

    dsp_PARAM();
    int nscoefs = dspDataTableFloat(noiseshaper, 3 * 6);
    int sine = dspGenerator_Sine(64);
    int dly = dspDelay_MicroSec_Max_Default(1000, 450);

    /* core 1: the global dither generator and what reads it */
    dsp_CORE();
    dsp_SERIAL(0x1234);
    dsp_TPDF_CALC(24);
    dsp_STORE(0);
    dsp_WHITE();
    dsp_STORE(1);
    dsp_LOAD_GAIN_Fixed(IN(0), 0.5);
    dsp_DITHER();
    dsp_SAT0DB();
    dsp_STORE(2);
    dsp_LOAD_GAIN_Fixed(IN(1), 0.5);
    dsp_DITHER_NS2(nscoefs);
    dsp_SAT0DB();
    dsp_STORE(3);

    /* core 2: core-local dither tables (DSP_TPDF), which also change the STORE mask of this core */
    dsp_CORE();
    dsp_TPDF(20);
    dsp_STORE(4);
    dsp_LOAD_GAIN_Fixed(IN(0), 0.7);
    dsp_SAT0DB_TPDF();
    dsp_STORE(5);
    dsp_LOAD_GAIN_Fixed(IN(1), 0.3);
    dsp_DITHER();
    dsp_SAT0DB();
    dsp_STORE(6);
    dsp_TPDF(16);
    dsp_LOAD_GAIN_Fixed(IN(1), 0.3);
    dsp_SAT0DB_TPDF_GAIN_Fixed(0.9);
    dsp_STORE(7);
    dsp_TPDF(0);                     /* back to the default dither through the local table */
    dsp_LOAD_GAIN_Fixed(IN(0), 0.2);
    dsp_DITHER_NS2(nscoefs);
    dsp_SAT0DB();
    dsp_STORE(16);

    /* core 3: meters and the DC blocker */
    dsp_CORE();
    dsp_LOAD_GAIN_Fixed(IN(0), 1.0);
    dsp_DCBLOCK(20);
    dsp_SAT0DB();
    dsp_STORE(17);
    dsp_LOAD(IN(0));
    dsp_RMS(10, 2);
    dsp_STORE(18);
    dsp_LOAD(IN(1));
    dsp_RMS(10, 0);
    dsp_STORE(19);
    dsp_LOAD(IN(0));
    dsp_COPYXY();
    dsp_MUL_FixedInt(3);
    dsp_DIV_FixedInt(4);
    dsp_PWRXY(10, 3);
    dsp_STORE(20);

    /* core 4: generators, clip, distribution, the 64-bit delays */
    dsp_CORE();
    dsp_CLRXY();
    dsp_DIRAC_Fixed(1000, 0.5);
    dsp_SAT0DB();
    dsp_STORE(21);
    dsp_SQUAREWAVE_Fixed(2000, 0.8);
    dsp_SAT0DB();
    dsp_STORE(22);
    dsp_DATA_TABLE(sine, 0.5, 3, 64);
    dsp_SAT0DB();
    dsp_STORE(23);
    dsp_LOAD_GAIN_Fixed(IN(0), 1.0);
    dsp_CLIP_Fixed(0.4);
    dsp_SAT0DB();
    dsp_STORE(24);
    dsp_LOAD(IN(1));
    dsp_DISTRIB(25, 64);
    dsp_LOAD_GAIN_Fixed(IN(0), 0.5);
    dsp_DELAY_1();
    dsp_SAT0DB();
    dsp_STORE(26);
    dsp_LOAD_GAIN_Fixed(IN(1), 0.5);
    dsp_DELAY_DP_FixedMicroSec(300);
    dsp_SAT0DB();
    dsp_STORE(27);
    dsp_LOAD_GAIN_Fixed(IN(0), 0.25);
    dsp_DELAY_DP(dly);
    dsp_SAT0DB();
    dsp_STORE(28);
    return dsp_END_OF_CODE();
}
