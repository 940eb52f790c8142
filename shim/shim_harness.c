/*
 * shim_harness.c -- drives shim/avdsp_plugin_b200.c the way alsa-lib would (open with an asound.conf node, init at a
 * rate, one transfer per period) so that the shim can be tested and timed without a sound card.  TEST INFRASTRUCTURE.
 *
 *   shim_harness <prog.bin> <rate> <format: s16|s24|s32> <period> <in.raw> <out.raw> [key=value ...] [--latency N]
 * in.raw: interleaved frames in <format>, as many channels as the plugin announces; out.raw: interleaved S32.
 * --latency N: after the pass over in.raw, time N more transfers of one period each and print p50 / p99 as JSON.
 */
#include <alsa/asoundlib.h>
#include <alsa/pcm_external.h>
#include <stdint.h>
#include <time.h>

int _snd_pcm_avdsp_open(snd_pcm_t **pcmp, const char *name, snd_config_t *root, snd_config_t *conf, snd_pcm_stream_t stream, int mode);

static int cmpd(const void *a, const void *b) { const double x = *(const double *)a, y = *(const double *)b; return x < y ? -1 : x > y; }

int main(int argc, char **argv) {
    if (argc < 7) { fprintf(stderr, "usage: %s prog.bin rate s16|s24|s32 period in.raw out.raw [key=value ...] [--latency N]\n", argv[0]); return 2; }
    const unsigned rate = (unsigned)atoi(argv[2]);
    const snd_pcm_format_t fmt = !strcmp(argv[3], "s16") ? SND_PCM_FORMAT_S16 : !strcmp(argv[3], "s24") ? SND_PCM_FORMAT_S24_3LE : SND_PCM_FORMAT_S32;
    const size_t bps = fmt == SND_PCM_FORMAT_S16 ? 2 : fmt == SND_PCM_FORMAT_S24_3LE ? 3 : 4;
    const unsigned long period = strtoul(argv[4], NULL, 10);
    int latencyN = 0;
    snd_config_t nodes[16], conf = {0};
    int nn = 0;
    memset(nodes, 0, sizeof nodes);
    nodes[nn].id = "slave"; nodes[nn].str = "null"; nn++;
    nodes[nn].id = "dspprog"; nodes[nn].str = argv[1]; nn++;
    for (int a = 7; a < argc && nn < 16; a++) {
        if (!strcmp(argv[a], "--latency") && a + 1 < argc) { latencyN = atoi(argv[++a]); continue; }
        char *eq = strchr(argv[a], '=');
        if (!eq) continue;
        *eq = 0;
        nodes[nn].id = argv[a];
        char *end; const long v = strtol(eq + 1, &end, 10);
        if (*end == 0 && end != eq + 1) { nodes[nn].is_num = 1; nodes[nn].num = v; } else nodes[nn].str = eq + 1;
        nn++;
    }
    for (int k = 0; k + 1 < nn; k++) nodes[k].next = &nodes[k + 1];
    conf.child = &nodes[0];

    snd_pcm_t *pcm = NULL;
    int rc = _snd_pcm_avdsp_open(&pcm, "avdsp", NULL, &conf, SND_PCM_STREAM_PLAYBACK, 0);
    if (rc < 0) { fprintf(stderr, "open failed: %d\n", rc); return 1; }
    snd_pcm_extplug_t *ext = (snd_pcm_extplug_t *)pcm;            /* the stub hands the extplug itself out as the pcm */
    ext->format = fmt; ext->rate = rate;
    if (ext->callback->init(ext) < 0) { fprintf(stderr, "init failed\n"); return 1; }
    const unsigned cin = ext->channels, cout = ext->slave_channels;

    FILE *fi = fopen(argv[5], "rb"), *fo = fopen(argv[6], "wb");
    if (!fi || !fo) { fprintf(stderr, "cannot open the PCM files\n"); return 1; }
    unsigned char *in = malloc(period * (cin ? cin : 1) * bps);
    int32_t *out = malloc(period * (cout ? cout : 1) * sizeof(int32_t));
    snd_pcm_channel_area_t sa = { in, 0, (unsigned)(cin * bps * 8) }, da = { out, 0, cout * 32 };
    size_t got;
    while ((got = fread(in, cin * bps, period, fi)) > 0) {
        memset(out, 0, period * cout * sizeof(int32_t));
        if (ext->callback->transfer(ext, &da, 0, &sa, 0, got) != (snd_pcm_sframes_t)got) { fprintf(stderr, "transfer failed\n"); return 1; }
        fwrite(out, sizeof(int32_t) * cout, got, fo);
    }
    fclose(fi); fclose(fo);
    if (latencyN > 0) {
        double *us = malloc(sizeof(double) * (size_t)latencyN);
        for (unsigned long k = 0; k < period * cin * bps; k++) in[k] = (unsigned char)(k * 37u + 11u);
        for (int k = -20; k < latencyN; k++) {                   /* 20 untimed warm-up periods */
            struct timespec a, b;
            clock_gettime(CLOCK_MONOTONIC, &a);
            ext->callback->transfer(ext, &da, 0, &sa, 0, period);
            clock_gettime(CLOCK_MONOTONIC, &b);
            if (k >= 0) us[k] = (b.tv_sec - a.tv_sec) * 1e6 + (b.tv_nsec - a.tv_nsec) * 1e-3;
        }
        qsort(us, (size_t)latencyN, sizeof(double), cmpd);
        printf("{\"period_frames\": %lu, \"rate\": %u, \"format\": \"%s\", \"channels_in\": %u, \"channels_out\": %u, \"transfers\": %d, "
               "\"p50_us\": %.1f, \"p99_us\": %.1f, \"min_us\": %.1f, \"period_us\": %.1f}\n", period, rate, argv[3], cin, cout, latencyN,
               us[latencyN / 2], us[(int)(latencyN * 0.99)], us[0], 1e6 * (double)period / rate);
    }
    ext->callback->close(ext);
    return 0;
}
