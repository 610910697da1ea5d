/* tools/hostprobe/null_engine.c -- DIAGNOSTIC: an instant "GPU" (every cmgpu_* call the host shim makes returns at once)
 * so that the host loop of csrc/host/shim_bench.c can be timed alone. Never part of the product. */
#include "cmgpu.h"
#include <stdlib.h>
#include <string.h>
struct cmgpu_ctx { unsigned channels, max_streams, slots, block_frames; size_t stride, slot_bytes; unsigned char *host; uint64_t launches; };
const char *cmgpu_last_error(void) { return ""; }
int cmgpu_device_count(void) { return 1; }
cmgpu_ctx_t *cmgpu_ctx_create(int device, unsigned channels, unsigned max_streams, unsigned ring_slots, unsigned block_frames, unsigned flags)
{ cmgpu_ctx_t *c = calloc(1, sizeof(*c)); c->channels=channels; c->max_streams=max_streams; c->slots=ring_slots; c->block_frames=block_frames;
  c->stride = ((size_t)block_frames*channels*2u+15u)&~(size_t)15u; c->slot_bytes=c->stride*max_streams; c->host = aligned_alloc(4096, ring_slots*c->slot_bytes); memset(c->host,0,ring_slots*c->slot_bytes); return c; }
void cmgpu_ctx_destroy(cmgpu_ctx_t *c){ free(c->host); free(c); }
size_t cmgpu_block_stride(const cmgpu_ctx_t *c){return c->stride;}
uint64_t cmgpu_launch_count(const cmgpu_ctx_t *c){return c->launches;}
void *cmgpu_host_slot(cmgpu_ctx_t *c, unsigned slot){return c->host+slot*c->slot_bytes;}
int cmgpu_stream_set_gain(cmgpu_ctx_t *c, unsigned s, unsigned n, uint16_t scale, const uint16_t *g){return 0;}
int cmgpu_slot_set_frames(cmgpu_ctx_t *c, unsigned slot, const uint32_t *f){return 0;}
int cmgpu_submit(cmgpu_ctx_t *c, unsigned slot, const void *h){return 0;}
int cmgpu_process(cmgpu_ctx_t *c, unsigned slot, unsigned flags){c->launches++;return 0;}
int cmgpu_fetch(cmgpu_ctx_t *c, unsigned slot, void *h){return 0;}
int cmgpu_slot_wait(cmgpu_ctx_t *c, unsigned slot){return 0;}
int cmgpu_sync(cmgpu_ctx_t *c){return 0;}
int cmgpu_meter_reset(cmgpu_ctx_t *c, unsigned f, unsigned n){return 0;}
int cmgpu_meter_results(cmgpu_ctx_t *c, unsigned first, unsigned count, uint32_t rate, int reset, unsigned flags, cmgpu_result_t *out, cmgpu_meter_state_t *x, int *rcs)
{ for (unsigned i=0;i<count;i++){ memset(&out[i],0,sizeof(out[i])); out[i].frames=1; rcs[i]=0;} return 0; }
int cmgpu_meter_result(cmgpu_ctx_t *c, unsigned s, uint32_t rate, cmgpu_result_t *out){memset(out,0,sizeof(*out)); return 0;}
