/* host/shim_internal.h -- shared plumbing of the coolmic_* object shim (see include/coolmic_b200_shim.h). */
#ifndef COOLMIC_B200_SHIM_INTERNAL_H
#define COOLMIC_B200_SHIM_INTERNAL_H

#include "../../../include/cmgpu.h"
#include "../../../include/coolmic_b200_shim.h"

#ifdef COOLMIC_B200_WITH_IGLOO
/* Integration build: libigloo objects, declared and allocated the way src/transform.c:60-75 does. */
typedef igloo_ro_base_t shim_base_t;
typedef igloo_ro_t shim_self_t;
#define SHIM_SELF(self, T)            igloo_RO_TO_TYPE(self, T)
#define SHIM_RO_NULL                  igloo_RO_NULL
#define SHIM_TYPE(T, cb)              igloo_RO_PUBLIC_TYPE(T, igloo_RO_TYPEDECL_FREE(cb))
#define SHIM_NEW(T, cb, name, assoc)  igloo_ro_new_raw(T, name, assoc)
#define shim_ref(o)   igloo_ro_ref(o)
#define shim_unref(o) igloo_ro_unref(o)
#else
/* Stand-alone object base: reference count + destructor, first member of every object. */
typedef struct shim_base {
    size_t refcount;
    void (*on_free)(void *self);
} shim_base_t;
typedef void *shim_self_t;

void *shim_alloc(size_t size, void (*on_free)(void *self));
#define SHIM_SELF(self, T)            ((T *)(self))
#define SHIM_RO_NULL                  NULL
#define SHIM_TYPE(T, cb)              typedef T shim_type_decl_unused_##T
#define SHIM_NEW(T, cb, name, assoc)  ((void)(name), (void)(assoc), (T *)shim_alloc(sizeof(T), cb))
#define shim_ref(o)   coolmic_b200_ref(o)
#define shim_unref(o) coolmic_b200_unref(o)
#endif

int  shim_device(void);
void shim_count_launches(uint64_t n);

#define SHIM_BLOCK_BYTES 8192u        /* the most the chain ever asks for at once (tee.c:91-97) */
#define SHIM_VU_BUFFER (2u * COOLMIC_B200_MAX_CHANNELS * 32u)      /* 1024, vumeter.c:48 */

struct coolmic_b200_batch;

/* coolmic_transform_t: stand-alone (private one-stream context) or a member of a batch */
struct coolmic_transform {
    shim_base_t base;
    coolmic_iohandle_t *io;
    unsigned char carry[2 * COOLMIC_B200_MAX_CHANNELS - 1];
    size_t carry_fill;
    uint_least32_t rate;
    unsigned int channels;
    /* setting kept on the host so that it survives until the context exists */
    uint16_t gain_scale;
    uint16_t gain[COOLMIC_B200_MAX_CHANNELS];
    int gain_dirty;
    cmgpu_ctx_t *ctx;                   /* stand-alone only */
    unsigned int block_frames;
    struct coolmic_b200_batch *batch;   /* batch member: the engine is shared ...          */
    unsigned int stream;                /* ... and this is the object's stream in it       */
};

/* coolmic_vumeter_t: stand-alone (meters what it pulls) or bound to a batch transform (fused) */
struct coolmic_vumeter {
    shim_base_t base;
    coolmic_iohandle_t *in;
    uint_least32_t rate;
    unsigned int channels;
    unsigned char buffer[SHIM_VU_BUFFER];
    size_t fill;
    cmgpu_ctx_t *ctx;                   /* stand-alone only */
    struct coolmic_b200_batch *batch;   /* fused: the device meters the transform's output  */
    unsigned int stream;
    uint64_t seen_bytes;                /* fused: bytes of metered output already reported by read() */
};

/* batch hooks used by the object code */
ssize_t shim_batch_read(struct coolmic_b200_batch *b, unsigned stream, void *cursor, void *buffer, size_t len);
void   *shim_batch_cursor_new(struct coolmic_b200_batch *b, unsigned stream);
void    shim_batch_cursor_free(struct coolmic_b200_batch *b, void *cursor);
size_t  shim_batch_cursor_unread(struct coolmic_b200_batch *b, void *cursor);
int     shim_batch_set_gain(struct coolmic_b200_batch *b, unsigned stream, uint16_t scale, const uint16_t *gain);
void    shim_batch_release(struct coolmic_b200_batch *b, unsigned stream);
ssize_t shim_batch_vumeter_read(struct coolmic_b200_batch *b, struct coolmic_vumeter *v, ssize_t maxlen);
int     shim_batch_vumeter_result(struct coolmic_b200_batch *b, struct coolmic_vumeter *v, coolmic_vumeter_result_t *out);
int     shim_batch_vumeter_reset(struct coolmic_b200_batch *b, struct coolmic_vumeter *v);

#endif
