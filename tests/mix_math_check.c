/* tests/mix_math_check.c -- TEST INFRASTRUCTURE (CPU): the integer identities the 8 -> 2 downmix kernel relies on
 * (libcoolmic-dsp_b200/csrc/cmgpu_mix.cuh, mix8_quot_raw / meter_batch4), restated in plain C and checked against
 * the specification  out = clamp16(trunc(sum_c x[c] * w[c] / scale))  (oracle_mix_process) on adversarial and
 * random inputs. The device code itself is checked on the GPU (tests/test_gpu_mix.py); this file checks that the
 * arithmetic it implements is exact, including the cases a random GPU test hardly ever hits: sums beyond 32 bits,
 * quotients at the clamp boundaries, scale 65535.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

static uint64_t rng = 0x9E3779B97F4A7C15ull;
static uint64_t next(void)
{
    uint64_t x = (rng += 0x9E3779B97F4A7C15ull);
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

static int spec(const int16_t *x, const uint16_t *w, uint16_t scale)
{
    int64_t n = 0, q;
    int c;
    for (c = 0; c < 8; c++)
        n += (int64_t)x[c] * w[c];
    q = n / scale;                                  /* C division truncates toward zero */
    return q > 32767 ? 32767 : q < -32768 ? -32768 : (int)q;
}

static unsigned ceil_log2(uint32_t v)
{
    unsigned b = 0;
    while ((1ull << b) < v)
        b++;
    return b;
}

/* mix8_quot_raw + the 16-bit saturation of cvt.pack.sat, with the magic / shift of cmgpu_stream_set_mix */
static int device_formula(const int16_t *x, const uint16_t *w, uint16_t scale)
{
    const unsigned l = ceil_log2(scale);
    const uint32_t shift = 31 + l;
    const uint32_t magic = (uint32_t)((((uint64_t)1 << shift) / scale) + 1);
    int32_t lo = 0, hi, ns, r;
    uint32_t a, q;
    int c;
    for (c = 0; c < 8; c++)
        lo += (int32_t)x[c] * (int32_t)(w[c] & 0xffu);          /* dp2a.lo: signed 16 x unsigned 8 */
    hi = lo >> 8;                                               /* arithmetic */
    for (c = 0; c < 8; c++)
        hi += (int32_t)x[c] * (int32_t)(w[c] >> 8);             /* dp2a.hi, seeded */
    if (hi > 0x7fffff) hi = 0x7fffff;
    if (hi < -0x7fffff) hi = -0x7fffff;
    ns = (int32_t)((uint32_t)hi * 256u + (uint32_t)(lo & 255)); /* |ns| <= 2^31 - 1: no wrap */
    a = (uint32_t)(ns < 0 ? -(int64_t)ns : ns);
    q = (uint32_t)(((uint64_t)a * magic) >> shift);
    r = (int32_t)q * ((ns >> 31) | 1);
    return r > 32767 ? 32767 : r < -32768 ? -32768 : r;
}

static int check(const int16_t *x, const uint16_t *w, uint16_t scale, const char *what)
{
    const int a = spec(x, w, scale), b = device_formula(x, w, scale);
    if (a != b) {
        printf("MISMATCH (%s): scale %u spec %d formula %d\n", what, scale, a, b);
        return 1;
    }
    return 0;
}

/* meter_batch4: two squares at a time in 32 bits, one 64-bit accumulate per batch; two 3-input maxima */
static int check_batch(void)
{
    int bad = 0, t, u;
    for (t = 0; t < 200000; t++) {
        int x[4];
        uint32_t k_ref = (uint32_t)next(), k_new, radd0 = 0xffffu - (uint32_t)(next() % 60000u), k[4], s01, s23;
        uint64_t p_ref = next() >> 8, p_new = p_ref;
        k_new = k_ref &= 0x7fffffffu;
        for (u = 0; u < 4; u++) {
            const unsigned r = (unsigned)(next() % 8u);
            x[u] = r == 0 ? -32768 : r == 1 ? 32767 : (int)(int16_t)next();
        }
        for (u = 0; u < 4; u++) {                           /* the per-frame code */
            const uint32_t key = ((uint32_t)abs(x[u]) << 16) + (radd0 - (uint32_t)u);
            if (key > k_ref)
                k_ref = key;
            p_ref += (uint64_t)((int64_t)x[u] * x[u]);
        }
        for (u = 0; u < 4; u++)
            k[u] = ((uint32_t)abs(x[u]) << 16) + (radd0 - (uint32_t)u);
        k_new = k_new > k[0] ? k_new : k[0]; k_new = k_new > k[1] ? k_new : k[1];
        k_new = k_new > k[2] ? k_new : k[2]; k_new = k_new > k[3] ? k_new : k[3];
        s01 = (uint32_t)(x[0] * x[0]) + (uint32_t)(x[1] * x[1]);
        s23 = (uint32_t)(x[2] * x[2]) + (uint32_t)(x[3] * x[3]);
        p_new += (uint64_t)s01 + (uint64_t)s23;
        bad |= k_ref != k_new || p_ref != p_new;
    }
    return bad;
}

int main(void)
{
    static const uint16_t scales[] = {1, 2, 3, 7, 255, 256, 257, 1000, 4096, 9999, 32767, 32768, 32769, 65534, 65535};
    static const int16_t xs[] = {-32768, -32767, -1, 0, 1, 32766, 32767};
    static const uint16_t ws[] = {0, 1, 255, 256, 257, 32768, 65280, 65535};
    int bad = 0;
    long cases = 0;
    unsigned si, i, j, c, t;
    int16_t x[8];
    uint16_t w[8];
    /* all channels equal: every combination of the corner samples and weights (sums up to 8 * 32768 * 65535 > 2^34) */
    for (si = 0; si < sizeof(scales) / sizeof(*scales); si++)
        for (i = 0; i < sizeof(xs) / sizeof(*xs); i++)
            for (j = 0; j < sizeof(ws) / sizeof(*ws); j++) {
                for (c = 0; c < 8; c++) { x[c] = xs[i]; w[c] = ws[j]; }
                bad |= check(x, w, scales[si], "uniform corners");
                for (c = 0; c < 8; c++) { x[c] = (c & 1) ? xs[i] : (int16_t)-xs[(i + 3) % 7]; w[c] = (c & 2) ? ws[j] : ws[(j + 5) % 8]; }
                bad |= check(x, w, scales[si], "mixed corners");
                cases += 2;
            }
    /* sums steered onto the clamp boundaries and onto the 2^31 neighbourhood, where the 32-bit path clamps H */
    for (t = 0; t < 3000000 && !bad; t++) {
        const uint16_t scale = (t & 7) == 0 ? 65535 : (t & 7) == 1 ? (uint16_t)(32768 + next() % 32768) : (uint16_t)(1 + next() % 65535);
        const unsigned mode = (unsigned)(next() % 4u);
        for (c = 0; c < 8; c++) {
            x[c] = (int16_t)next();
            w[c] = (uint16_t)next();
        }
        if (mode == 1) {                                    /* large sums of one sign */
            const int neg = (int)(next() & 1);
            for (c = 0; c < 8; c++) {
                x[c] = (int16_t)(neg ? -(int)(20000 + next() % 12769) : (int)(20000 + next() % 12768));
                w[c] = (uint16_t)(next() % 65536);
            }
        } else if (mode == 2) {                             /* one dominant channel: n close to +-32768 * scale */
            for (c = 1; c < 8; c++) { x[c] = (int16_t)(next() % 5) - 2; w[c] = (uint16_t)(next() % 4); }
            x[0] = (int16_t)((next() & 1) ? -32768 + (int)(next() % 3) : 32767 - (int)(next() % 3));
            w[0] = (uint16_t)(scale - (uint16_t)(next() % 3));
        } else if (mode == 3) {                             /* n around +-2^31 */
            const int neg = (int)(next() & 1);
            for (c = 0; c < 8; c++) { x[c] = 0; w[c] = 0; }
            x[0] = neg ? -32768 : 32767; w[0] = 65535;      /* +-2.147e9 */
            x[1] = (int16_t)((int)(next() % 2001) - 1000); w[1] = (uint16_t)(next() % 65536);
            x[2] = (int16_t)((int)(next() % 201) - 100); w[2] = (uint16_t)(next() % 300);
        }
        bad |= check(x, w, scale, "random");
        cases++;
    }
    bad |= check_batch();
    printf("mix_math_check: %ld quotient cases, 200000 meter batches: %s\n", cases, bad ? "FAILED" : "ok");
    return bad;
}
