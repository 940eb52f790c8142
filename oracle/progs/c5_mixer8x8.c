/*
 * Config C5 (BASELINE.json configs[4]): 8-in x 8-out matrix mixer + per-channel
 * sample delays + gain/dither.  Our own DSP program on the unchanged reference
 * encoder API.  One core: TPDF_CALC(24), then per output
 *    LOAD_MUX(8 inputs) -> SAT0DB_TPDF_GAIN -> DELAY(<=5 ms) -> STORE.
 * usage: dspcreate -dspprog c5_mixer8x8.so -binfile x.bin -dspformat 2 -fsmin 192000 -fsmax 192000
 */
#include <stdlib.h>
#include <string.h>
#include "dsp_encoder.h"

int dspProg(int argc, char **argv) {
    int maxus = 5000;
    for (int i = 0; i < argc; i++)
        if (strcmp(argv[i], "-maxus") == 0 && i + 1 < argc) maxus = strtol(argv[++i], NULL, 10);

    dsp_CORE();
    dsp_TPDF_CALC(24);

    dsp_PARAM();
    int mux[8], dly[8];
    for (int o = 0; o < 8; o++) {
        mux[o] = dspLoadMux_Inputs(8);
        for (int i = 0; i < 8; i++) {
            /* diagonal-dominant mixing matrix with alternating-sign cross terms */
            float g = (i == o) ? 0.70f : (((i + o) & 1) ? -0.04f : 0.05f) * (1.0f + 0.1f * i);
            dspLoadMux_Data(8 + i, g);
        }
    }
    for (int o = 0; o < 8; o++)
        dly[o] = dspDelay_MicroSec_Max_Default(maxus, 100 + 650 * o);   /* 0.1 .. 4.65 ms */

    for (int o = 0; o < 8; o++) {
        dsp_LOAD_MUX(mux[o]);
        dsp_SAT0DB_TPDF_GAIN_Fixed(0.9f - 0.05f * o);
        dsp_DELAY(dly[o]);
        dsp_STORE(o);
    }
    return dsp_END_OF_CODE();
}
