/* oracle/vorbis_shim/vorbis/vorbisenc.h -- TEST INFRASTRUCTURE: see codec.h in this directory. */
#ifndef ORACLE_VORBIS_SHIM_VORBISENC_H
#define ORACLE_VORBIS_SHIM_VORBISENC_H
#include "codec.h"
int vorbis_encode_init_vbr(vorbis_info *vi, long channels, long rate, float base_quality);
#endif
