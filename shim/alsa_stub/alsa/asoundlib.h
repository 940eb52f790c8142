/*
 * Minimal stand-in for <alsa/asoundlib.h> + <alsa/pcm_external.h>: just enough of the external-plugin API to COMPILE and
 * DRIVE shim/avdsp_plugin_b200.c in a container without alsa-lib (no alsa headers in the image).
 * TEST INFRASTRUCTURE.  Names and signatures follow alsa-lib's public headers; the bodies are trivial (a configuration
 * node is a key with a string / integer value or a list of children).  With real alsa-lib installed, build the shim
 * against it instead (shim/Makefile: ALSA=1).
 */
#ifndef AVDSP_ALSA_STUB_H_
#define AVDSP_ALSA_STUB_H_
#include <errno.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned long snd_pcm_uframes_t;
typedef long snd_pcm_sframes_t;
typedef struct snd_pcm snd_pcm_t;
typedef enum { SND_PCM_STREAM_PLAYBACK = 0, SND_PCM_STREAM_CAPTURE } snd_pcm_stream_t;
typedef enum { SND_PCM_FORMAT_S16 = 2, SND_PCM_FORMAT_S32 = 10, SND_PCM_FORMAT_S24_3LE = 32 } snd_pcm_format_t;
typedef struct { void *addr; unsigned int first, step; } snd_pcm_channel_area_t;

/* configuration tree */
typedef struct snd_config {
    const char *id; const char *str; long num; int is_num;
    struct snd_config *child, *next;
} snd_config_t;
typedef snd_config_t *snd_config_iterator_t;
#define snd_config_for_each(i, nxt, node) for (i = (node)->child, nxt = i ? i->next : NULL; i; i = nxt, nxt = i ? i->next : NULL)
static inline snd_config_t *snd_config_iterator_entry(snd_config_iterator_t i) { return i; }
static inline int snd_config_get_id(const snd_config_t *n, const char **id) { *id = n->id; return 0; }
static inline int snd_config_get_string(const snd_config_t *n, const char **v) { if (n->is_num || !n->str) return -EINVAL; *v = n->str; return 0; }
static inline int snd_config_get_integer(const snd_config_t *n, long *v) { if (!n->is_num) return -EINVAL; *v = n->num; return 0; }
#define SNDERR(...) do { fprintf(stderr, "ALSA stub: " __VA_ARGS__); fprintf(stderr, "\n"); } while (0)

/* external filter plugin */
#define SND_PCM_EXTPLUG_VERSION ((1 << 16) | (0 << 8) | 2)
enum { SND_PCM_EXTPLUG_HW_FORMAT = 0, SND_PCM_EXTPLUG_HW_CHANNELS = 1 };
typedef struct snd_pcm_extplug snd_pcm_extplug_t;
typedef struct snd_pcm_extplug_callback {
    snd_pcm_sframes_t (*transfer)(snd_pcm_extplug_t *ext, const snd_pcm_channel_area_t *dst_areas, snd_pcm_uframes_t dst_offset,
                                  const snd_pcm_channel_area_t *src_areas, snd_pcm_uframes_t src_offset, snd_pcm_uframes_t size);
    int (*close)(snd_pcm_extplug_t *ext);
    int (*hw_params)(snd_pcm_extplug_t *ext, void *params);
    int (*hw_free)(snd_pcm_extplug_t *ext);
    void (*dump)(snd_pcm_extplug_t *ext, void *out);
    int (*init)(snd_pcm_extplug_t *ext);
} snd_pcm_extplug_callback_t;
struct snd_pcm_extplug {
    unsigned int version; const char *name; const snd_pcm_extplug_callback_t *callback; void *private_data; snd_pcm_t *pcm;
    snd_pcm_stream_t stream; snd_pcm_format_t format; int subformat; unsigned int channels; unsigned int rate;
    snd_pcm_format_t slave_format; int slave_subformat; unsigned int slave_channels;
};
static inline int snd_pcm_extplug_create(snd_pcm_extplug_t *ext, const char *name, snd_config_t *root, snd_config_t *sconf,
                                         snd_pcm_stream_t stream, int mode) {
    (void)name; (void)root; (void)sconf; (void)mode; ext->stream = stream; ext->pcm = (snd_pcm_t *)ext; return 0;
}
static inline int snd_pcm_extplug_set_param(snd_pcm_extplug_t *ext, int type, unsigned int v) { if (type == SND_PCM_EXTPLUG_HW_CHANNELS) ext->channels = v; return 0; }
static inline int snd_pcm_extplug_set_slave_param(snd_pcm_extplug_t *ext, int type, unsigned int v) {
    if (type == SND_PCM_EXTPLUG_HW_CHANNELS) ext->slave_channels = v; else ext->slave_format = (snd_pcm_format_t)v; return 0;
}
static inline int snd_pcm_extplug_set_param_list(snd_pcm_extplug_t *ext, int type, unsigned int n, const unsigned int *list) {
    (void)ext; (void)type; (void)n; (void)list; return 0;
}
#define SND_PCM_PLUGIN_DEFINE_FUNC(plugin) \
    int _snd_pcm_##plugin##_open(snd_pcm_t **pcmp, const char *name, snd_config_t *root, snd_config_t *conf, snd_pcm_stream_t stream, int mode)
#define SND_PCM_PLUGIN_SYMBOL(plugin)
#endif
