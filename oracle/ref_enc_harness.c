/* oracle/ref_enc_harness.c -- TEST INFRASTRUCTURE.
 *
 * Runs the sample-format stage of the reference's OWN src/enc_vorbis.c (unmodified, compiled from
 * /root/reference by oracle/Makefile target `refenc`): a coolmic_enc_t is set up by hand the way
 * enc.c would, fed from a memory iohandle, and its codec callback `process` is called until the
 * input is exhausted. Every call pulls 1,024 bytes, de-interleaves them and writes
 * `sample / 32768.f` into the buffer vorbis_analysis_buffer() hands out (enc_vorbis.c:76-122).
 * The libvorbis / libogg entry points below are the stand-in of oracle/vorbis_shim: no codec, no
 * arithmetic -- they give the reference code planes to write into and count what it wrote.
 * this pins the product's CMGPU_PLANAR output against reference object code.
 */
#include <stdlib.h>
#include <string.h>

#include "types_private.h"
#include <coolmic-dsp/coolmic-dsp.h>
#include <coolmic-dsp/enc.h>
#include "enc_private.h"

/* ---- capture area shared by the stand-in functions ---- */
static float *g_plane[16];
static float *g_cursor[16];
static size_t g_capacity, g_frames;
static unsigned g_channels;

void vorbis_info_init(vorbis_info *vi) { memset(vi, 0, sizeof(*vi)); }
void vorbis_info_clear(vorbis_info *vi) { (void)vi; }
int vorbis_encode_init_vbr(vorbis_info *vi, long channels, long rate, float q) { vi->channels = (int)channels; vi->rate = rate; (void)q; return 0; }
void vorbis_comment_init(vorbis_comment *vc) { memset(vc, 0, sizeof(*vc)); }
void vorbis_comment_add_tag(vorbis_comment *vc, const char *t, const char *c) { (void)vc; (void)t; (void)c; }
void vorbis_comment_clear(vorbis_comment *vc) { (void)vc; }
int vorbis_analysis_init(vorbis_dsp_state *v, vorbis_info *vi) { memset(v, 0, sizeof(*v)); v->vi = vi; return 0; }
int vorbis_block_init(vorbis_dsp_state *v, vorbis_block *vb) { vb->vd = v; vb->opaque = NULL; return 0; }
int vorbis_block_clear(vorbis_block *vb) { (void)vb; return 0; }
void vorbis_dsp_clear(vorbis_dsp_state *v) { (void)v; }
int vorbis_analysis_headerout(vorbis_dsp_state *v, vorbis_comment *vc, ogg_packet *a, ogg_packet *b, ogg_packet *c)
{
    (void)v; (void)vc;
    memset(a, 0, sizeof(*a)); memset(b, 0, sizeof(*b)); memset(c, 0, sizeof(*c));
    return 0;
}
float **vorbis_analysis_buffer(vorbis_dsp_state *v, int vals)
{
    unsigned c;
    (void)v;
    if (g_frames + (size_t)vals > g_capacity)
        return NULL;                                   /* (never: the harness sizes the planes for the whole input) */
    for (c = 0; c < g_channels; c++)
        g_cursor[c] = g_plane[c] + g_frames;
    return g_cursor;
}
int vorbis_analysis_wrote(vorbis_dsp_state *v, int vals) { (void)v; g_frames += (size_t)vals; return 0; }
int vorbis_analysis_blockout(vorbis_dsp_state *v, vorbis_block *vb) { (void)v; (void)vb; return 0; }      /* "need more data" */
int vorbis_analysis(vorbis_block *vb, ogg_packet *op) { (void)vb; (void)op; return 0; }
int vorbis_bitrate_addblock(vorbis_block *vb) { (void)vb; return 0; }
int vorbis_bitrate_flushpacket(vorbis_dsp_state *vd, ogg_packet *op) { (void)vd; (void)op; return 0; }
int ogg_stream_packetin(ogg_stream_state *os, ogg_packet *op) { (void)os; (void)op; return 0; }
int coolmic_metadata_add_to_vorbis_comment(coolmic_metadata_t *self, vorbis_comment *vc) { (void)self; (void)vc; return 0; }

typedef struct memsrc {
    const char *data;
    size_t len, pos, framesize;
} memsrc_t;

/* hands out whole frames only, as the transform in front of the encoder does (transform.c:133-137) */
static ssize_t memsrc_read(void *userdata, void *buffer, size_t len)
{
    memsrc_t *m = userdata;
    size_t n = m->len - m->pos;
    len -= len % m->framesize;
    if (n > len)
        n = len;
    memcpy(buffer, m->data + m->pos, n);
    m->pos += n;
    return (ssize_t)n;
}
static int memsrc_eof(void *userdata)
{
    memsrc_t *m = userdata;
    return m->pos >= m->len;
}

/* interleaved S16 `in` (in_bytes, whole frames) -> planes[channels][frames] float, written by the reference's
 * __vorbis_read_data through its own `process` callback. Returns the frames it reports, or < 0. */
long refenc_vorbis_planes(const void *in, size_t in_bytes, unsigned channels, float *planes, size_t plane_stride)
{
    coolmic_enc_t *self;
    coolmic_iohandle_t *src;
    memsrc_t mem;
    unsigned c;
    int guard;

    if (!channels || channels > 16 || in_bytes % (2u * channels))
        return -1;
    self = calloc(1, sizeof(*self));
    if (!self)
        return -2;
    mem.data = in;
    mem.len = in_bytes;
    mem.pos = 0;
    mem.framesize = 2u * channels;
    src = coolmic_iohandle_new("memsrc", igloo_RO_NULL, &mem, NULL, memsrc_read, memsrc_eof);
    self->state = STATE_RUNNING;
    self->rate = 48000;
    self->channels = channels;
    self->in = src;
    self->quality = 0.1f;
    self->cb = __coolmic_enc_cb_vorbis;
    g_channels = channels;
    g_capacity = plane_stride;
    g_frames = 0;
    for (c = 0; c < channels; c++)
        g_plane[c] = planes + (size_t)c * plane_stride;
    if (self->cb.start(self) != 0)
        return -3;
    /* a process() call loops over __vorbis_read_data until that reports something other than 0: EOF (-1), or
     * "nothing now" (-2) when iohandle_read stopped at a short read */
    for (guard = 0; guard < 1000000 && self->state != STATE_EOF; guard++)
        self->cb.process(self);
    self->cb.stop(self);
    igloo_ro_unref(src);
    free(self);
    return (long)g_frames;
}
