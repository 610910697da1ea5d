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
