/* host/shim_object.c -- reference-counted object base and the iohandle boundary object.
 *
 * iohandle semantics follow reference src/iohandle.c:54-113: an object wrapping read/eof/free
 * callbacks; read() re-invokes the callback until the request is filled, the callback reports
 * "nothing now" (0) or fails, and reports progress in preference to a late error.
 */
#include "shim_internal.h"

#include <stdlib.h>

static int g_device = -1;
static uint64_t g_launches;

#ifndef COOLMIC_B200_WITH_IGLOO
void *shim_alloc(size_t size, void (*on_free)(void *self))
{
    shim_base_t *b = calloc(1, size);
    if (!b)
        return NULL;
    b->refcount = 1;
    b->on_free = on_free;
    return b;
}

int coolmic_b200_ref(coolmic_b200_ro_t object)
{
    shim_base_t *b = object;
    if (!b)
        return COOLMIC_ERROR_FAULT;
    __atomic_add_fetch(&b->refcount, 1, __ATOMIC_RELAXED);
    return COOLMIC_ERROR_NONE;
}

int coolmic_b200_unref(coolmic_b200_ro_t object)
{
    shim_base_t *b = object;
    if (!b)
        return COOLMIC_ERROR_FAULT;
    if (__atomic_sub_fetch(&b->refcount, 1, __ATOMIC_ACQ_REL) == 0) {
        if (b->on_free)
            b->on_free(b);
        free(b);
    }
    return COOLMIC_ERROR_NONE;
}
#endif /* !COOLMIC_B200_WITH_IGLOO */

int shim_device(void)
{
    if (g_device < 0) {
        const char *e = getenv("COOLMIC_B200_DEVICE");
        g_device = e ? atoi(e) : 0;
    }
    return g_device;
}

int coolmic_b200_set_device(int device)
{
    if (device < 0 || device >= cmgpu_device_count())
        return COOLMIC_ERROR_INVAL;
    g_device = device;
    return COOLMIC_ERROR_NONE;
}

void shim_count_launches(uint64_t n) { __atomic_add_fetch(&g_launches, n, __ATOMIC_RELAXED); }
uint64_t coolmic_b200_shim_launches(void) { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

/* ---- iohandle (stand-alone build only; the integration build keeps the reference's iohandle.c) */
#ifndef COOLMIC_B200_WITH_IGLOO
struct coolmic_iohandle {
    shim_base_t base;
    void *userdata;
    int (*free_cb)(void *);
    ssize_t (*read_cb)(void *, void *, size_t);
    int (*eof_cb)(void *);
};

static void iohandle_destroy(void *self)
{
    coolmic_iohandle_t *h = self;
    if (h->free_cb)
        h->free_cb(h->userdata);
}

coolmic_iohandle_t *coolmic_iohandle_new(const char *name, coolmic_b200_ro_t associated, void *userdata,
                                         int (*free_cb)(void *), ssize_t (*read_cb)(void *, void *, size_t),
                                         int (*eof_cb)(void *))
{
    coolmic_iohandle_t *h;
    (void)name, (void)associated;
    if (!read_cb)                       /* iohandle.c:58-60: a handle without read makes no sense */
        return NULL;
    h = shim_alloc(sizeof(*h), iohandle_destroy);
    if (!h)
        return NULL;
    h->userdata = userdata;
    h->free_cb = free_cb;
    h->read_cb = read_cb;
    h->eof_cb = eof_cb;
    return h;
}

ssize_t coolmic_iohandle_read(coolmic_iohandle_t *self, void *buffer, size_t len)
{
    size_t got = 0;
    if (!self || !buffer)
        return COOLMIC_ERROR_FAULT;
    if (!len)
        return 0;
    while (got < len) {
        ssize_t r = self->read_cb(self->userdata, (char *)buffer + got, len - got);
        if (r < 0)
            return got ? (ssize_t)got : r;
        if (r == 0)
            break;
        got += (size_t)r;
    }
    return (ssize_t)got;
}

int coolmic_iohandle_eof(coolmic_iohandle_t *self)
{
    if (!self)
        return COOLMIC_ERROR_FAULT;
    return self->eof_cb ? self->eof_cb(self->userdata) : 0;
}
#endif /* !COOLMIC_B200_WITH_IGLOO */
