/* tests/stub/san_main.c -- TEST INFRASTRUCTURE: a fixed scenario through the product's host shim on
 * the CPU stub engine, for runs under -fsanitize=address,undefined and -fsanitize=thread
 * (tests/test_shim_host_logic.py builds and runs it; exit code 0 = clean and results as expected).
 */
#include "../../include/coolmic_b200_shim.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

long shimh_batch(const void *in, size_t n_streams, size_t bytes_per_stream, unsigned rate, unsigned channels,
                 const uint16_t *scale, const uint16_t *gain, size_t src_chunk, unsigned block_frames,
                 unsigned result_every_ticks, size_t pull, void *out, size_t *out_bytes,
                 void *results, size_t cap, size_t *n_results, int *fanout_mismatch);
long shimh_batch_ring(const void *in, size_t n_streams, size_t bytes_per_stream, unsigned rate, unsigned channels,
                      const uint16_t *scale, const uint16_t *gain, size_t src_chunk, unsigned block_frames,
                      unsigned slots, unsigned threads, size_t pull, void *out, size_t *out_bytes,
                      void *results, int *flags);
int shimh_null_checks(void);

int shim_bench_copy_selftest(void);
void cmgpu_stub_fail_next_process(int n);
long shimh_transform(const void *in, size_t in_bytes, unsigned rate, unsigned channels, int set_gain, unsigned gain_n,
                     unsigned scale, const uint16_t *gain, size_t src_chunk, size_t pull, void *out, size_t out_cap, int *gain_rc);
long shimh_chain(const void *in, size_t in_bytes, unsigned rate, unsigned channels, int set_gain, unsigned gain_n,
                 unsigned scale, const uint16_t *gain, size_t src_chunk, long maxlen, unsigned result_every,
                 void *results, size_t results_cap);
long oracle_run_transform(const void *in, size_t in_bytes, unsigned channels, int set_gain, unsigned gain_n, unsigned scale,
                          const uint16_t *gain, size_t src_chunk, size_t pull, void *out, size_t out_cap, int *gain_rc);

/* the stand-alone objects: whole-frame reads with the partial-frame carry, odd chunkings and pull sizes, the gain
 * setter's four cases (transform.c:126-165,195-222) -- bytes against the oracle port, memory under the sanitizers */
static int standalone_objects_scenario(void)
{
    enum { BYTES = 9001 };
    static const unsigned chans[] = {1, 2, 3, 8, 16};
    static const size_t chunks[] = {0, 1, 3, 7, 100}, pulls[] = {1024, 33, 5, 8192};
    unsigned char *in = malloc(BYTES), *a = malloc(BYTES + 64), *b = malloc(BYTES + 64);
    void *results = calloc(64, 256);
    uint16_t gain[16];
    int bad = 0;
    unsigned i, t = 0;
    for (i = 0; i < BYTES; i++)
        in[i] = (unsigned char)rand();
    for (i = 0; i < 16; i++)
        gain[i] = (uint16_t)(300 * i + 77);
    for (i = 0; i < sizeof(chans) / sizeof(*chans); i++) {
        unsigned gi;
        for (gi = 0; gi < 4; gi++, t++) {
            const unsigned ch = chans[i];
            const unsigned gn = gi == 0 ? ch : gi == 1 ? 1 : gi == 2 ? 2 : 0;      /* copy, broadcast, mono average or INVAL, off */
            const size_t chunk = chunks[t % 5], pull = pulls[t % 4];
            int rc_a = 99, rc_b = 99;
            const long na = shimh_transform(in, BYTES, 48000, ch, 1, gn, 1000 + 13 * t, gain, chunk, pull, a, BYTES + 64, &rc_a);
            const long nb = oracle_run_transform(in, BYTES, ch, 1, gn, 1000 + 13 * t, gain, chunk, pull, b, BYTES + 64, &rc_b);
            bad |= na != nb || rc_a != rc_b || (na > 0 && memcmp(a, b, (size_t)na) != 0);
            bad |= shimh_chain(in, BYTES, 48000, ch, 1, gn, 1000 + 13 * t, gain, chunk, t % 2 ? -1 : 100, 3, results, 64) <= 0;
        }
    }
    free(in); free(a); free(b); free(results);
    return bad;
}

/* a tick that fails after its pull (a CUDA error in the real engine): nothing of it is exposed, and the
 * NEXT tick still hands the pull workers a new job instead of waiting for them forever */
typedef struct { const unsigned char *p; size_t left; } src_t;
static ssize_t src_read(void *u, void *buf, size_t len)
{
    src_t *m = u;
    if (len > m->left)
        len = m->left;
    memcpy(buf, m->p, len);
    m->p += len;
    m->left -= len;
    return (ssize_t)len;
}
static int src_eof(void *u) { return ((src_t *)u)->left == 0; }

static int failed_tick_scenario(unsigned threads)
{
    enum { N = 9, BLOCK = 64, BYTES = 4 * BLOCK * 3 };
    static unsigned char pcm[N][BYTES];
    unsigned char got[BYTES];
    coolmic_b200_batch_t *b = coolmic_b200_batch_new_ring(0, 2, N, BLOCK, 2, threads);
    coolmic_transform_t *tr[N];
    coolmic_iohandle_t *rd[N];
    src_t src[N];
    unsigned s, i;
    int bad = !b, rc;
    if (bad)
        return 1;
    for (s = 0; s < N; s++) {
        coolmic_iohandle_t *h;
        for (i = 0; i < BYTES; i++)
            pcm[s][i] = (unsigned char)(s * 31 + i * 7);
        src[s].p = pcm[s];
        src[s].left = BYTES;
        tr[s] = coolmic_b200_batch_transform_new(b, "tr", NULL, 48000);          /* no gain: output = input */
        h = coolmic_iohandle_new("src", NULL, &src[s], NULL, src_read, src_eof);
        coolmic_transform_attach_iohandle(tr[s], h);
        coolmic_b200_unref(h);
        rd[s] = coolmic_transform_get_iohandle(tr[s]);
    }
    cmgpu_stub_fail_next_process(1);
    rc = coolmic_b200_batch_tick(b);
    bad |= rc != COOLMIC_ERROR_GENERIC;
    for (s = 0; s < N; s++)
        bad |= coolmic_iohandle_read(rd[s], got, sizeof(got)) != 0;              /* the failed tick exposed nothing */
    bad |= coolmic_b200_batch_pending(b) != 0;
    rc = coolmic_b200_batch_tick(b);                                              /* hung here before the fix */
    bad |= rc != N * BLOCK;
    for (s = 0; s < N; s++) {
        /* the block the failed tick had pulled is gone (a read error loses data in the reference too); the next one arrives intact */
        bad |= coolmic_iohandle_read(rd[s], got, sizeof(got)) != 4 * BLOCK;
        bad |= memcmp(got, pcm[s] + 4 * BLOCK, 4 * BLOCK) != 0;
        coolmic_b200_unref(rd[s]);
        coolmic_b200_unref(tr[s]);
    }
    coolmic_b200_unref(b);
    return bad;
}

int main(void)
{
    enum { N = 37, CH = 3, BYTES = 2 * CH * 1777 + 5 };
    unsigned char *in = malloc((size_t)N * BYTES), *out = malloc((size_t)N * BYTES), *out2 = malloc((size_t)N * BYTES);
    size_t out_bytes[N], out_bytes2[N], n_results[N];
    uint16_t scale[N], gain[N * CH];
    void *results = calloc((size_t)N * 64, 256);
    int flags = 0, mism = 0, bad = 0;
    unsigned s, i;
    long ticks;
    double secs = 0;
    uint64_t check = 0;

    srand(7);
    for (i = 0; i < (unsigned)N * BYTES; i++)
        in[i] = (unsigned char)rand();
    for (s = 0; s < N; s++) {
        scale[s] = (uint16_t)(s % 5 ? 1000 + 13 * s : 0);
        for (i = 0; i < CH; i++)
            gain[s * CH + i] = (uint16_t)(700 + 97 * s + i);
    }
    bad |= shimh_null_checks() != 0;
    bad |= standalone_objects_scenario();
    bad |= failed_tick_scenario(1);
    bad |= failed_tick_scenario(4);
    /* the synchronous batch and the ring batch (3 slots, 4 pull threads) must produce the same bytes */
    ticks = shimh_batch(in, N, BYTES, 48000, CH, scale, gain, 11, 200, 2, 333, out, out_bytes, results, 64, n_results, &mism);
    bad |= ticks <= 0 || mism != 0;
    ticks = shimh_batch_ring(in, N, BYTES, 48000, CH, scale, gain, 11, 200, 3, 4, 333, out2, out_bytes2, results, &flags);
    bad |= ticks <= 0 || flags != 0;
    for (s = 0; s < N; s++)
        bad |= out_bytes[s] != out_bytes2[s] || memcmp(out + (size_t)s * BYTES, out2 + (size_t)s * BYTES, out_bytes[s]) != 0;
    /* the measurement driver: producer and consumer threads overlap */
    {
        enum { STREAMS = 64, BLOCK = 500, TICKS = 9 };
        const unsigned bps = 2 * 2 * BLOCK * 3;
        unsigned char *pcm = malloc((size_t)STREAMS * bps);
        for (i = 0; i < STREAMS * bps; i++)
            pcm[i] = (unsigned char)rand();
        bad |= shim_bench_copy_selftest();      /* the capture copy with non-temporal stores == memcpy */
        bad |= coolmic_b200_bench_objects(0, 2, STREAMS, BLOCK, TICKS, 4, 6, bps, pcm, &secs, &check) != 0;
        bad |= check != (uint64_t)STREAMS * BLOCK * TICKS;
        bad |= coolmic_b200_bench_objects(0, 2, STREAMS, BLOCK, TICKS, 1, 3, bps, pcm, &secs, &check) != 0;
        bad |= coolmic_b200_bench_objects(0, 2, STREAMS, BLOCK, TICKS, 2, 1, bps, pcm, &secs, &check) != 0;
        free(pcm);
    }
    free(in); free(out); free(out2); free(results);
    printf("san_main: %s\n", bad ? "FAILED" : "ok");
    return bad;
}
