// cmgpu_post.cu -- what sits either side of the fused tick on the device:
//
//   * taking meter rows for MANY streams in one stream-ordered step (copy + conditional reset, and
//     optionally the dB finaliser of vumeter.c:198-212 in fp64 on the device), the building block of
//     cmgpu_meter_result, cmgpu_meter_results and cmgpu_gather_results;
//   * meter colours (util.c:59-139) for every stream in one kernel;
//   * the tone / noise source that fills ring slots without a host upload (snddev_sine.c:118-150 as a
//     cyclic read of a run-time table).
//
// None of this is on the hot path's roofline: tables of a few MB per reporting interval.
#include "cmgpu_ctx.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <thread>

using cmgpu::fail;

namespace cmgpu {

// One thread per stream: copy the row out, clear it when frames were metered (what result() does per
// object, vumeter.c:198-199,214-215 -- a row without frames is all zero anyway), optionally finalise.
static __global__ void meter_take(unsigned long long *rows, unsigned row_u64, unsigned channels, unsigned count,
                                  unsigned long long *out_rows, int reset, cmgpu_result_t *out_results, uint32_t rate)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count)
        return;
    unsigned long long *row = rows + (size_t)i * row_u64;
    unsigned long long *dst = out_rows + (size_t)i * row_u64;
    uint64_t local[2 * CMGPU_MAX_CHANNELS + 2];
    for (unsigned j = 0; j < row_u64; j++) {
        local[j] = row[j];
        dst[j] = local[j];
    }
    if (reset && local[2 * channels]) {
        for (unsigned j = 0; j < row_u64; j++)
            row[j] = 0;
    }
    if (out_results) {
        cmgpu_meter_state_t st;
        decode_row(local, channels, &st);
        cmgpu_result_t res;
        if (finalise_state(&st, rate, channels, &res) != CMGPU_OK) {
            for (unsigned k = 0; k < sizeof(res) / 8; k++)
                reinterpret_cast<uint64_t *>(&res)[k] = 0;
        }
        out_results[i] = res;
    }
}

// ---- colours (util.c:59-139) ------------------------------------------------------------------
__device__ __forceinline__ uint32_t unit_to_byte(double x)
{
    // clamp to [0, 1], scale, truncate (util.c:30-44)
    if (x >= 1.)
        x = 1.;
    else if (x <= 0.)
        x = 0.;
    const uint32_t v = (uint32_t)(x * 255.);
    return v > 255u ? 255u : v;
}

__device__ uint32_t ahsv2argb(double alpha, double hue, double saturation, double value)
{
    // util.c:59-106: the fractional part is taken of `hue` itself, as the reference does
    const double kPi = 3.14159265358979323846;
    const int sector = (int)(hue / (kPi / 3.));
    const double f = hue - (double)sector;
    const double p = value * (1. - saturation);
    const double q = value * (1. - saturation * f);
    const double t = value * (1. - saturation * (1. - f));
    double r = 0., g = 0., b = 0.;
    switch (sector) {
    case 0: case 6: r = value; g = t;     b = p;     break;
    case 1:         r = q;     g = value; b = p;     break;
    case 2:         r = p;     g = value; b = t;     break;
    case 3:         r = p;     g = q;     b = value; break;
    case 4:         r = t;     g = p;     b = value; break;
    case 5:         r = value; g = p;     b = q;     break;
    default:        break;
    }
    return (unit_to_byte(alpha) << 24) + (unit_to_byte(r) << 16) + (unit_to_byte(g) << 8) + unit_to_byte(b);
}

__device__ double power2hue(double power)
{
    // util.c:110-122 ("default" profile): green below -20 dB, red at 0 dB, sin^2 ramp in between
    const double kPi = 3.14159265358979323846;
    if (power < -20.)
        return kPi * 2. / 3.;
    if (power >= 0)
        return 0;
    const double s = sin(kPi * power / 40.);
    return s * s * kPi * 2. / 3.;          // pow(s, 2.) is correctly rounded = s * s
}

__device__ double peak2hue(int peak)
{
    // util.c:126-139
    const double kPi = 3.14159265358979323846;
    if (peak == -32768 || peak == 32767)
        return 0.;
    if (peak < -30000 || peak > 30000)
        return 0.43;
    if (peak < -28000 || peak > 28000)
        return 1.;
    return kPi * 2. / 3.;
}

static __global__ void meter_colors(const unsigned long long *rows, unsigned row_u64, unsigned channels, unsigned count,
                                    double alpha, double saturation, double value, cmgpu_colors_t *out)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count)
        return;
    uint64_t local[2 * CMGPU_MAX_CHANNELS + 2];
    for (unsigned j = 0; j < row_u64; j++)
        local[j] = rows[(size_t)i * row_u64 + j];
    cmgpu_meter_state_t st;
    decode_row(local, channels, &st);
    cmgpu_result_t res;
    cmgpu_colors_t col;
    for (unsigned k = 0; k < sizeof(col) / 8; k++)
        reinterpret_cast<uint64_t *>(&col)[k] = 0;
    if (finalise_state(&st, 0, channels, &res) == CMGPU_OK) {
        col.global_power_hue = power2hue(res.global_power);
        col.global_power_argb = ahsv2argb(alpha, col.global_power_hue, saturation, value);
        col.global_peak_argb = ahsv2argb(alpha, peak2hue(res.global_peak), saturation, value);
        for (unsigned c = 0; c < channels; c++) {
            col.channel_power_hue[c] = power2hue(res.channel_power[c]);
            col.channel_power_argb[c] = ahsv2argb(alpha, col.channel_power_hue[c], saturation, value);
            col.channel_peak_argb[c] = ahsv2argb(alpha, peak2hue(res.channel_peak[c]), saturation, value);
        }
    }
    out[i] = col;
}

// ---- tone / noise source ----------------------------------------------------------------------
// One thread per 16-byte vector of the slot; the table (<= 4,096 samples) is read through L1/L2.
static __global__ void tone_fill(uint8_t *slot, const uint32_t *frames, unsigned n_streams, unsigned block_frames,
                                 unsigned channels, size_t stride_bytes, const int16_t *period, unsigned n,
                                 uint64_t first_frame, unsigned first_stream, unsigned stream_step, unsigned channel_step)
{
    const unsigned vec_per_block = (unsigned)(stride_bytes / 16);
    const uint64_t total = (uint64_t)n_streams * vec_per_block;
    for (uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; v < total; v += (uint64_t)gridDim.x * blockDim.x) {
        const unsigned s = (unsigned)(v / vec_per_block);
        const unsigned k = (unsigned)(v - (uint64_t)s * vec_per_block);
        const unsigned nfr = frames ? min(frames[s], block_frames) : block_frames;
        const uint64_t valid = (uint64_t)nfr * channels;
        uint64_t i0 = (uint64_t)k * 8u;
        unsigned f = (unsigned)(i0 / channels);
        unsigned c = (unsigned)(i0 - (uint64_t)f * channels);
        const uint64_t sbase = (first_frame + (uint64_t)stream_step * (first_stream + s)) % n;
        uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
        for (int j = 0; j < 8; j++) {
            int16_t x = 0;
            if (i0 + j < valid)
                x = period[(sbase + f + (uint64_t)channel_step * c) % n];
            w[j >> 1] |= (uint32_t)(uint16_t)x << ((j & 1) * 16);
            if (++c == channels) {
                c = 0;
                f++;
            }
        }
        *reinterpret_cast<uint4 *>(slot + (size_t)s * stride_bytes + (size_t)k * 16) = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

__host__ __device__ inline uint64_t splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

static __global__ void noise_fill(uint8_t *slot, const uint32_t *frames, unsigned n_streams, unsigned block_frames,
                                  unsigned channels, size_t stride_bytes, uint64_t first_frame, unsigned first_stream,
                                  uint64_t seed, unsigned every, unsigned phase)
{
    const unsigned vec_per_block = (unsigned)(stride_bytes / 16);
    const uint64_t total = (uint64_t)n_streams * vec_per_block;
    for (uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; v < total; v += (uint64_t)gridDim.x * blockDim.x) {
        const unsigned s = (unsigned)(v / vec_per_block);
        const uint64_t gs = (uint64_t)first_stream + s;
        if (gs % every != phase)
            continue;
        const unsigned k = (unsigned)(v - (uint64_t)s * vec_per_block);
        const unsigned nfr = frames ? min(frames[s], block_frames) : block_frames;
        const uint64_t valid = (uint64_t)nfr * channels;
        uint64_t i0 = (uint64_t)k * 8u;
        unsigned f = (unsigned)(i0 / channels);
        unsigned c = (unsigned)(i0 - (uint64_t)f * channels);
        uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
        for (int j = 0; j < 8; j++) {
            int16_t x = 0;
            if (i0 + j < valid)
                x = (int16_t)splitmix64(seed ^ (gs << 40) ^ ((first_frame + f) << 4) ^ (uint64_t)c);
            w[j >> 1] |= (uint32_t)(uint16_t)x << ((j & 1) * 16);
            if (++c == channels) {
                c = 0;
                f++;
            }
        }
        *reinterpret_cast<uint4 *>(slot + (size_t)s * stride_bytes + (size_t)k * 16) = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

// ---- host side ------------------------------------------------------------------------------------

int ensure_take_buffers_locked(cmgpu_ctx *c)
{
    const unsigned row = c->row_u64 > c->row_in_u64 ? c->row_u64 : c->row_in_u64;
    const size_t bytes = sizeof(uint64_t) * row * c->max_streams;
    if (!c->d_take)
        CU(cudaMalloc(&c->d_take, bytes));
    if (!c->h_take)
        CU(cudaMallocHost(&c->h_take, bytes));
    return CMGPU_OK;
}

// Queues, on the compute stream: rows [first, first+count) -> d_take (+ conditional reset, + device
// finaliser into d_results when asked). Nothing is waited for.
int take_rows_locked(cmgpu_ctx *c, unsigned first, unsigned count, int reset, bool device_db, uint32_t rate)
{
    int rc = ensure_take_buffers_locked(c);
    if (rc)
        return rc;
    if (device_db && !c->d_results)
        CU(cudaMalloc(&c->d_results, sizeof(cmgpu_result_t) * c->max_streams));
    const unsigned C = c->out_channels ? c->out_channels : c->channels;
    meter_take<<<(count + 127) / 128, 128, 0, c->cmp()>>>(c->d_meters + (size_t)first * c->row_u64, c->row_u64, C, count,
                                                          c->d_take, reset, device_db ? c->d_results : nullptr, rate);
    CU(cudaGetLastError());
    c->last_first = ~0u;            // something other than a tick is now last on the compute stream
    c->chain_open = false;
    if (reset && first == 0 && count == c->max_streams && !c->out_channels) {
        // every window starts afresh: bring the position base back to zero (the keys keep 46 - pbits
        // bits of tick number; a window must not span more ticks than that)
        CU(cudaMemsetAsync(c->d_tick, 0, sizeof(unsigned long long), c->cmp()));
        c->pending_ticks = 0;
    }
    return CMGPU_OK;
}

void finalise_rows(const uint64_t *rows, size_t count, unsigned row_u64, unsigned channels, uint32_t rate,
                   cmgpu_result_t *results, cmgpu_meter_state_t *states, int *rcs)
{
    auto work = [=](size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; i++) {
            cmgpu_meter_state_t st;
            decode_row(rows + i * row_u64, channels, &st);
            if (states)
                states[i] = st;
            int r = st.frames ? CMGPU_OK : CMGPU_ERR_INVAL;
            if (results) {
                r = finalise_state(&st, rate, channels, results + i);
                if (r != CMGPU_OK)
                    memset(results + i, 0, sizeof(*results));
            }
            if (rcs)
                rcs[i] = r;
        }
    };
    unsigned n_threads = 1;
    if (count >= 8192 && results) {
        n_threads = std::thread::hardware_concurrency();
        n_threads = n_threads > 8 ? 8 : (n_threads ? n_threads : 1);
    }
    if (n_threads <= 1) {
        work(0, count);
        return;
    }
    std::vector<std::thread> pool;
    for (unsigned t = 1; t < n_threads; t++)
        pool.emplace_back(work, count * t / n_threads, count * (t + 1) / n_threads);
    work(0, count / n_threads);
    for (auto &th : pool)
        th.join();
}

}  // namespace cmgpu

extern "C" {

int cmgpu_meter_results(cmgpu_ctx_t *c, unsigned first, unsigned count, uint32_t rate, int reset, unsigned flags,
                        cmgpu_result_t *results, cmgpu_meter_state_t *states, int *rcs)
{
    CMGPU_TRACE("cmgpu_meter_results");
    if (!c)
        return fail(CMGPU_ERR_FAULT, "NULL context");
    if ((uint64_t)first + count > c->max_streams)
        return fail(CMGPU_ERR_INVAL, "stream range out of bounds");
    if (!count)
        return CMGPU_OK;
    const bool device_db = (flags & CMGPU_RESULTS_DEVICE_DB) != 0 && results;
    const unsigned C = c->out_channels ? c->out_channels : c->channels;
    std::lock_guard<std::mutex> lk(c->mu);
    CU(cudaSetDevice(c->device));
    int rc = cmgpu::take_rows_locked(c, first, count, reset, device_db, rate);
    if (rc)
        return rc;
    CU(cudaMemcpyAsync(c->h_take, c->d_take, sizeof(uint64_t) * c->row_u64 * count, cudaMemcpyDeviceToHost, c->cmp()));
    if (device_db)
        CU(cudaMemcpyAsync(results, c->d_results, sizeof(cmgpu_result_t) * count, cudaMemcpyDeviceToHost, c->cmp()));
    CU(cudaStreamSynchronize(c->cmp()));
    if (device_db) {
        // the dB values came from the device; only the integer state and the return codes are derived here
        cmgpu::finalise_rows(c->h_take, count, c->row_u64, C, rate, nullptr, states, rcs);
    } else {
        cmgpu::finalise_rows(c->h_take, count, c->row_u64, C, rate, results, states, rcs);
    }
    return CMGPU_OK;
}

int cmgpu_meter_result(cmgpu_ctx_t *c, unsigned stream, uint32_t rate, cmgpu_result_t *out)
{
    // coolmic_vumeter_result (vumeter.c:189-218): copy out and reset as ONE step; no frames -> INVAL, no reset
    if (!c || !out)
        return fail(CMGPU_ERR_FAULT, "NULL argument");
    cmgpu_result_t res;
    int r = CMGPU_OK;
    int rc = cmgpu_meter_results(c, stream, 1, rate, 1, 0, &res, nullptr, &r);
    if (rc)
        return rc;
    if (r != CMGPU_OK)
        return r;
    *out = res;
    return CMGPU_OK;
}

int cmgpu_meter_colors(cmgpu_ctx_t *c, unsigned first, unsigned count, double alpha, double saturation, double value,
                       cmgpu_colors_t *out)
{
    if (!c || !out)
        return fail(CMGPU_ERR_FAULT, "NULL argument");
    if ((uint64_t)first + count > c->max_streams)
        return fail(CMGPU_ERR_INVAL, "stream range out of bounds");
    if (!count)
        return CMGPU_OK;
    const unsigned C = c->out_channels ? c->out_channels : c->channels;
    std::lock_guard<std::mutex> lk(c->mu);
    CU(cudaSetDevice(c->device));
    if (!c->d_colors)
        CU(cudaMalloc(&c->d_colors, sizeof(cmgpu_colors_t) * c->max_streams));
    cmgpu::meter_colors<<<(count + 127) / 128, 128, 0, c->cmp()>>>(c->d_meters + (size_t)first * c->row_u64, c->row_u64, C,
                                                                   count, alpha, saturation, value,
                                                                   static_cast<cmgpu_colors_t *>(c->d_colors));
    CU(cudaGetLastError());
    c->last_first = ~0u;
    c->chain_open = false;
    CU(cudaMemcpyAsync(out, c->d_colors, sizeof(cmgpu_colors_t) * count, cudaMemcpyDeviceToHost, c->cmp()));
    CU(cudaStreamSynchronize(c->cmp()));
    return CMGPU_OK;
}

int cmgpu_tone_set_table(cmgpu_ctx_t *c, const int16_t *period, unsigned n)
{
    if (!c || !period)
        return fail(CMGPU_ERR_FAULT, "NULL argument");
    if (!n || n > 4096)
        return fail(CMGPU_ERR_INVAL, "tone table of 1..4096 samples");
    std::lock_guard<std::mutex> lk(c->mu);
    CU(cudaSetDevice(c->device));
    if (!c->d_tone)
        CU(cudaMalloc(&c->d_tone, sizeof(int16_t) * 4096));
    // pageable source: staged by the runtime before the call returns
    CU(cudaMemcpyAsync(c->d_tone, period, sizeof(int16_t) * n, cudaMemcpyHostToDevice, c->cmp()));
    c->tone_len = n;
    return CMGPU_OK;
}

static int fill_prologue_locked(cmgpu_ctx *c, unsigned slot)
{
    // like an upload: the slot must not be overwritten while a download of it is in flight, and the
    // fill runs on the compute stream, so ticks before and after it are ordered by the stream itself
    if (c->down_pending[slot]) {
        CU(cudaStreamWaitEvent(c->cmp(), c->ev_down[slot], 0));
        c->down_pending[slot] = 0;
    }
    if (c->up_pending[slot]) {
        CU(cudaStreamWaitEvent(c->cmp(), c->ev_up[slot], 0));
        c->up_pending[slot] = 0;
    }
    c->last_first = ~0u;
    c->chain_open = false;
    return CMGPU_OK;
}

int cmgpu_tone_fill(cmgpu_ctx_t *c, unsigned slot, uint64_t first_frame, unsigned first_stream, unsigned stream_step,
                    unsigned channel_step)
{
    if (!c || slot >= c->slots)
        return fail(c ? CMGPU_ERR_INVAL : CMGPU_ERR_FAULT, "bad context or slot");
    std::lock_guard<std::mutex> lk(c->mu);
    if (!c->d_tone || !c->tone_len)
        return fail(CMGPU_ERR_INVAL, "no tone table (cmgpu_tone_set_table)");
    if (!c->active)
        return CMGPU_OK;
    CU(cudaSetDevice(c->device));
    int rc = fill_prologue_locked(c, slot);
    if (rc)
        return rc;
    const uint64_t vecs = (uint64_t)c->active * (c->stride / 16);
    const unsigned grid = (unsigned)std::min<uint64_t>((vecs + 255) / 256, (uint64_t)c->num_sms * 16);
    cmgpu::tone_fill<<<grid, 256, 0, c->cmp()>>>(c->d_in + (size_t)slot * c->slot_bytes,
                                                 c->has_frames[slot] ? c->d_frames + (size_t)slot * c->max_streams : nullptr,
                                                 c->active, c->block_frames, c->channels, c->stride, c->d_tone, c->tone_len,
                                                 first_frame, first_stream, stream_step, channel_step);
    CU(cudaGetLastError());
    c->cmp_unrecorded[slot] = 1;
    c->slot_dirty[slot] = 1;
    return CMGPU_OK;
}

int cmgpu_noise_fill(cmgpu_ctx_t *c, unsigned slot, uint64_t first_frame, unsigned first_stream, uint64_t seed,
                     unsigned every, unsigned phase)
{
    if (!c || slot >= c->slots)
        return fail(c ? CMGPU_ERR_INVAL : CMGPU_ERR_FAULT, "bad context or slot");
    if (!every || phase >= every)
        return fail(CMGPU_ERR_INVAL, "every >= 1 and phase < every");
    std::lock_guard<std::mutex> lk(c->mu);
    if (!c->active)
        return CMGPU_OK;
    CU(cudaSetDevice(c->device));
    int rc = fill_prologue_locked(c, slot);
    if (rc)
        return rc;
    const uint64_t vecs = (uint64_t)c->active * (c->stride / 16);
    const unsigned grid = (unsigned)std::min<uint64_t>((vecs + 255) / 256, (uint64_t)c->num_sms * 16);
    cmgpu::noise_fill<<<grid, 256, 0, c->cmp()>>>(c->d_in + (size_t)slot * c->slot_bytes,
                                                  c->has_frames[slot] ? c->d_frames + (size_t)slot * c->max_streams : nullptr,
                                                  c->active, c->block_frames, c->channels, c->stride, first_frame,
                                                  first_stream, seed, every, phase);
    CU(cudaGetLastError());
    c->cmp_unrecorded[slot] = 1;
    c->slot_dirty[slot] = 1;
    return CMGPU_OK;
}

int cmgpu_link_probe(int device, size_t bytes, unsigned reps, int write_combined, float *h2d_gbs, float *d2h_gbs,
                     float *both_gbs)
{
    if (!bytes || !reps)
        return fail(CMGPU_ERR_INVAL, "bytes and reps must be non-zero");
    CU(cudaSetDevice(device));
    char *h_in = nullptr, *h_out = nullptr, *d_in = nullptr, *d_out = nullptr;
    cudaStream_t up = nullptr, down = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr;
    int rc = CMGPU_OK;
    auto run = [&]() -> int {
        CU(cudaHostAlloc(&h_in, bytes, write_combined ? cudaHostAllocWriteCombined : cudaHostAllocDefault));
        CU(cudaMallocHost(&h_out, bytes));
        CU(cudaMalloc(&d_in, bytes));
        CU(cudaMalloc(&d_out, bytes));
        memset(h_in, 1, bytes);
        CU(cudaMemset(d_out, 2, bytes));
        CU(cudaStreamCreateWithFlags(&up, cudaStreamNonBlocking));
        CU(cudaStreamCreateWithFlags(&down, cudaStreamNonBlocking));
        CU(cudaEventCreate(&e0));
        CU(cudaEventCreate(&e1));
        CU(cudaEventCreate(&e2));
        float ms = 0.f;
        const double gb = (double)bytes * reps / 1e9;
        if (h2d_gbs) {
            CU(cudaMemcpyAsync(d_in, h_in, bytes, cudaMemcpyHostToDevice, up));          // warm-up
            CU(cudaEventRecord(e0, up));
            for (unsigned r = 0; r < reps; r++)
                CU(cudaMemcpyAsync(d_in, h_in, bytes, cudaMemcpyHostToDevice, up));
            CU(cudaEventRecord(e1, up));
            CU(cudaEventSynchronize(e1));
            CU(cudaEventElapsedTime(&ms, e0, e1));
            *h2d_gbs = (float)(gb / (ms * 1e-3));
        }
        if (d2h_gbs) {
            CU(cudaMemcpyAsync(h_out, d_out, bytes, cudaMemcpyDeviceToHost, down));
            CU(cudaEventRecord(e0, down));
            for (unsigned r = 0; r < reps; r++)
                CU(cudaMemcpyAsync(h_out, d_out, bytes, cudaMemcpyDeviceToHost, down));
            CU(cudaEventRecord(e1, down));
            CU(cudaEventSynchronize(e1));
            CU(cudaEventElapsedTime(&ms, e0, e1));
            *d2h_gbs = (float)(gb / (ms * 1e-3));
        }
        if (both_gbs) {
            CU(cudaDeviceSynchronize());
            CU(cudaEventRecord(e0, up));
            CU(cudaStreamWaitEvent(down, e0, 0));
            for (unsigned r = 0; r < reps; r++) {
                CU(cudaMemcpyAsync(d_in, h_in, bytes, cudaMemcpyHostToDevice, up));
                CU(cudaMemcpyAsync(h_out, d_out, bytes, cudaMemcpyDeviceToHost, down));
            }
            CU(cudaEventRecord(e2, down));
            CU(cudaStreamWaitEvent(up, e2, 0));
            CU(cudaEventRecord(e1, up));
            CU(cudaEventSynchronize(e1));
            CU(cudaEventElapsedTime(&ms, e0, e1));
            *both_gbs = (float)(gb / (ms * 1e-3));
        }
        return CMGPU_OK;
    };
    rc = run();
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (e2) cudaEventDestroy(e2);
    if (up) cudaStreamDestroy(up);
    if (down) cudaStreamDestroy(down);
    cudaFree(d_in);
    cudaFree(d_out);
    if (h_in) cudaFreeHost(h_in);
    if (h_out) cudaFreeHost(h_out);
    return rc;
}

}  // extern "C"
