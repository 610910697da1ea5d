/* oracle/igloo_shim/igloo/ro.h -- TEST INFRASTRUCTURE, not product code.
 *
 * Minimal stand-in for the part of libigloo's reference-counted object ("RO")
 * surface that the reference hot-path files use. libigloo
 * is an external, un-vendored dependency of libcoolmic-dsp (linked as -ligloo,
 * reference src/Makefile.inc:28) that contributes only object plumbing, no
 * arithmetic. This shim exists solely so oracle/Makefile can compile the
 * reference's own .c files unmodified into oracle/_ref/.
 *
 * Semantics provided:
 *   - igloo_ro_t is an untyped pointer; every object starts with igloo_ro_base_t.
 *   - igloo_ro_new_raw(T, name, assoc): zeroed sizeof(T) allocation, refcount 1.
 *   - igloo_ro_ref / igloo_ro_unref: return 0 on success, non-zero on NULL;
 *     unref at zero runs the type's free callback then releases the memory.
 */
#ifndef ORACLE_IGLOO_SHIM_RO_H
#define ORACLE_IGLOO_SHIM_RO_H

#include <stddef.h>
#include <stdlib.h>

typedef void *igloo_ro_t;

typedef struct igloo_shim_type {
    size_t      length;
    const char *name;
    void      (*on_free)(igloo_ro_t self);
} igloo_ro_type_t;

typedef struct igloo_shim_base {
    const igloo_ro_type_t *type;
    size_t                 refcount;
} igloo_ro_base_t;

#define igloo_RO_NULL            ((igloo_ro_t)0)
#define igloo_RO_TO_TYPE(x, T)   ((T *)(x))
#define igloo_RO_IS_NULL(x)      ((x) == igloo_RO_NULL)

/* type-description helpers: each expands to designated initialisers */
#define igloo_RO_TYPEDECL_FREE(cb)  .on_free = (cb)
#define igloo_RO_TYPEDECL_NEW(cb)   /* constructors via igloo_ro_new() are not used by the oracle */

#define igloo_RO__DESCRIPTOR(T)     igloo_shim_typedesc__##T
#define igloo_RO_PUBLIC_TYPE(T, ...) \
    const igloo_ro_type_t igloo_RO__DESCRIPTOR(T) = { .length = sizeof(T), .name = #T, __VA_ARGS__ }
#define igloo_RO_PRIVATE_TYPE(T, ...) \
    static const igloo_ro_type_t igloo_RO__DESCRIPTOR(T) = { .length = sizeof(T), .name = #T, __VA_ARGS__ }

static inline igloo_ro_t igloo_shim_alloc(const igloo_ro_type_t *type)
{
    igloo_ro_base_t *base = calloc(1, type->length);
    if (!base)
        return igloo_RO_NULL;
    base->type = type;
    base->refcount = 1;
    return base;
}

#define igloo_ro_new_raw(T, name, associated) \
    ((void)(name), (void)(associated), (T *)igloo_shim_alloc(&igloo_RO__DESCRIPTOR(T)))

static inline int igloo_ro_ref(igloo_ro_t self)
{
    igloo_ro_base_t *base = self;
    if (!base)
        return -1;
    base->refcount++;
    return 0;
}

static inline int igloo_ro_unref(igloo_ro_t self)
{
    igloo_ro_base_t *base = self;
    if (!base)
        return -1;
    if (--base->refcount)
        return 0;
    if (base->type->on_free)
        base->type->on_free(self);
    free(base);
    return 0;
}

#endif
