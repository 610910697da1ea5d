/* oracle/vorbis_shim/vorbis/codec.h -- TEST INFRASTRUCTURE.
 *
 * Header-only stand-in for the parts of libogg / libvorbis that reference src/enc_vorbis.c and
 * src/enc_private.h name, so that the reference's OWN enc_vorbis.c compiles unmodified here (the
 * libraries themselves are not installed) and its sample-format stage -- the de-interleave and
 * `sample / 32768.f` of enc_vorbis.c:108-117, SURVEY.md 8f N2 -- can be run as the oracle for the
 * product's float-plane output. The stand-in contributes NO arithmetic: the types are opaque
 * blobs, and the functions are implemented by oracle/ref_enc_harness.c to hand the reference code
 * a buffer to write its floats into and to record how many frames it says it wrote.
 */
#ifndef ORACLE_VORBIS_SHIM_CODEC_H
#define ORACLE_VORBIS_SHIM_CODEC_H

#include <stdint.h>

typedef int64_t ogg_int64_t;
typedef struct { unsigned char *packet; long bytes; long b_o_s; long e_o_s; ogg_int64_t granulepos; ogg_int64_t packetno; } ogg_packet;
typedef struct { unsigned char *header; long header_len; unsigned char *body; long body_len; } ogg_page;
typedef struct { int serialno; void *opaque; } ogg_stream_state;

typedef struct { int version; int channels; long rate; void *codec_setup; } vorbis_info;
typedef struct { char **user_comments; int *comment_lengths; int comments; char *vendor; } vorbis_comment;
typedef struct { int analysisp; vorbis_info *vi; void *opaque; } vorbis_dsp_state;
typedef struct { vorbis_dsp_state *vd; void *opaque; } vorbis_block;

void vorbis_info_init(vorbis_info *vi);
void vorbis_info_clear(vorbis_info *vi);
void vorbis_comment_init(vorbis_comment *vc);
void vorbis_comment_add_tag(vorbis_comment *vc, const char *tag, const char *contents);
void vorbis_comment_clear(vorbis_comment *vc);
int  vorbis_analysis_init(vorbis_dsp_state *v, vorbis_info *vi);
int  vorbis_block_init(vorbis_dsp_state *v, vorbis_block *vb);
int  vorbis_block_clear(vorbis_block *vb);
void vorbis_dsp_clear(vorbis_dsp_state *v);
int  vorbis_analysis_headerout(vorbis_dsp_state *v, vorbis_comment *vc, ogg_packet *op, ogg_packet *op_comm, ogg_packet *op_code);
float **vorbis_analysis_buffer(vorbis_dsp_state *v, int vals);
int  vorbis_analysis_wrote(vorbis_dsp_state *v, int vals);
int  vorbis_analysis_blockout(vorbis_dsp_state *v, vorbis_block *vb);
int  vorbis_analysis(vorbis_block *vb, ogg_packet *op);
int  vorbis_bitrate_addblock(vorbis_block *vb);
int  vorbis_bitrate_flushpacket(vorbis_dsp_state *vd, ogg_packet *op);
int  ogg_stream_packetin(ogg_stream_state *os, ogg_packet *op);

#endif
