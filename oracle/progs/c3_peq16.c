/*
 * Config C3 (BASELINE.json configs[2]): 16-section parametric EQ per channel.
 * Our own DSP program; it only calls the unchanged reference encoder API
 * (module_avdsp/encoder/dsp_encoder.h, dsp_filters.h).  One DSP_CORE per channel,
 * ALSA I/O convention of module_avdsp/linux/avdsp_plugin.c:29-32 (in = 8+ch, out = ch).
 * usage: dspcreate -dspprog c3_peq16.so -binfile x.bin -dspformat 3 -fsmin 48000 -fsmax 48000 [-ch N]
 */
#include <stdlib.h>
#include <string.h>
#include "dsp_encoder.h"

static const struct { float f, q, db; } band[16] = {
    {   40, 0.9f,  3.0f }, {   63, 1.4f, -2.5f }, {  100, 2.0f,  1.5f }, {  160, 1.0f, -4.0f },
    {  250, 3.0f,  2.0f }, {  400, 1.2f, -1.0f }, {  630, 4.0f,  3.5f }, { 1000, 0.7f, -3.0f },
    { 1600, 2.5f,  1.0f }, { 2500, 1.8f, -2.0f }, { 4000, 5.0f,  4.0f }, { 6300, 1.1f, -1.5f },
    { 8000, 2.2f,  2.5f }, {10000, 3.3f, -3.5f }, {12500, 1.6f,  0.5f }, {17000, 0.8f, -0.5f },
};

int dspProg(int argc, char **argv) {
    int nch = 2;
    for (int i = 0; i < argc; i++)
        if (strcmp(argv[i], "-ch") == 0 && i + 1 < argc) nch = strtol(argv[++i], NULL, 10);
    if (nch < 1) nch = 1;
    if (nch > 8) nch = 8;
    for (int ch = 0; ch < nch; ch++) {
        dsp_CORE();
        dsp_PARAM();
        int eq = dspBiquad_Sections(16);
        for (int i = 0; i < 16; i++)   /* each channel gets slightly different centre frequencies */
            dsp_Filter2ndOrder(FPEAK, band[i].f * (1.0f + 0.03f * ch), band[i].q, dB2gain(band[i].db));
        dsp_LOAD_GAIN_Fixed(8 + ch, 0.5);
        dsp_BIQUADS(eq);
        dsp_SAT0DB();
        dsp_STORE(ch);
    }
    return dsp_END_OF_CODE();
}
