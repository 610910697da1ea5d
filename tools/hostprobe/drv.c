/* tools/hostprobe/drv.c -- DIAGNOSTIC: the object-API bench driver (csrc/host/shim_bench.c) linked against
 * null_engine.c, an "instant GPU" whose cmgpu_* calls do nothing. What is left is the host side of the loop
 * alone: pulling every member's input into the ring slots and copying every stream's output out through
 * its iohandle -- the two host copies per sample the iohandle contract costs. Its rate is the ceiling of
 * bench.py's e2e.object_api on this host whatever the GPU does (bench.py reports it as host_loop_alone).
 * usage: hostloop THREADS SLOTS [STREAMS BLOCK_FRAMES TICKS]   -> prints best-of-4 seconds */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

int coolmic_b200_bench_objects(int device, unsigned int channels, unsigned int streams, unsigned int block_frames,
                               unsigned int n_ticks, unsigned int ring_slots, unsigned int threads,
                               unsigned int bytes_per_stream, const void *pcm, double *seconds, uint64_t *frames_metered);

int main(int argc, char **argv)
{
    const unsigned threads = argc > 1 ? (unsigned)atoi(argv[1]) : 8, slots = argc > 2 ? (unsigned)atoi(argv[2]) : 4;
    const unsigned streams = argc > 3 ? (unsigned)atoi(argv[3]) : 1024, block = argc > 4 ? (unsigned)atoi(argv[4]) : 12000;
    const unsigned ticks = argc > 5 ? (unsigned)atoi(argv[5]) : 40;
    const size_t bps = (size_t)block * 4 * 4;          /* stereo, four ticks' worth per source, like bench.py */
    unsigned char *pcm = malloc(streams * bps);
    double best = 1e9, s = 0;
    uint64_t fm = 0;
    int r, rc = 0;
    if (!pcm)
        return 1;
    for (size_t i = 0; i < streams * bps; i++)
        pcm[i] = (unsigned char)(i * 7);
    for (r = 0; r < 4 && !rc; r++) {
        rc = coolmic_b200_bench_objects(0, 2, streams, block, ticks, slots, threads, (unsigned)bps, pcm, &s, &fm);
        if (!rc && s < best)
            best = s;
    }
    if (rc) {
        printf("error %d\n", rc);
        return 1;
    }
    printf("seconds %.6f samples %.0f gbs_each_way %.2f threads %u slots %u\n", best,
           (double)streams * 2 * block * ticks, (double)streams * block * 4 * ticks / best / 1e9, threads, slots);
    return 0;
}
