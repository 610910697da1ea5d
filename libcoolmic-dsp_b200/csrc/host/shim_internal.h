/* host/shim_internal.h -- shared plumbing of the coolmic_* object shim (see include/coolmic_b200_shim.h). */
#ifndef COOLMIC_B200_SHIM_INTERNAL_H
#define COOLMIC_B200_SHIM_INTERNAL_H

#include "../../../include/cmgpu.h"
#include "../../../include/coolmic_b200_shim.h"

#ifdef COOLMIC_B200_WITH_IGLOO
#define shim_ref(o)   igloo_ro_ref(o)
#define shim_unref(o) igloo_ro_unref(o)
#error "integration build: declare the three types with igloo_RO_PUBLIC_TYPE and allocate with igloo_ro_new_raw (INTEGRATION.md)"
#else
/* Stand-alone object base: reference count + destructor, first member of every object. */
typedef struct shim_base {
    size_t refcount;
    void (*on_free)(void *self);
} shim_base_t;

void *shim_alloc(size_t size, void (*on_free)(void *self));
#define shim_ref(o)   coolmic_b200_ref(o)
#define shim_unref(o) coolmic_b200_unref(o)
#endif

int  shim_device(void);
void shim_count_launches(uint64_t n);

#endif
