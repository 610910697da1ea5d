// cmgpu.cu -- C-ABI batch engine (include/cmgpu.h) on top of the fused kernels.
//
// Host side of the hot path: owns the device ring, the per-stream gain recipes, the
// per-stream meter rows, three CUDA streams (upload / compute / download) and the events
// that order a slot's submit -> tick -> fetch. No PyTorch, no CPU fallback: every data-path
// entry point ends in a CUDA call and fails with CMGPU_ERR_GENERIC if that call fails.
#include "cmgpu_kernels.cuh"
#include "cmgpu_mix.cuh"
#include "cmgpu_tma.cuh"
#include "cmgpu_span.cuh"

#include "cmgpu_ctx.h"

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <mutex>
#include <new>
#include <algorithm>
#include <vector>

using cmgpu::MixArgs;
using cmgpu::TickArgs;
using cmgpu::fail;

namespace cmgpu {

static thread_local char g_err[512] = "";

int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

}  // namespace cmgpu

namespace {

// ---- the gain recipe (see GainRow in cmgpu_kernels.cuh, proof in DESIGN.md) ---------------
struct RecipeHost {
    uint32_t mw, addm, mul;
};

RecipeHost make_recipe(uint16_t g, uint16_t d)
{
    // d != 0 here. ratio = g/d in [0, 65535].
    unsigned pre = 0;
    if (g >= d) {
        // 2^t <= g/d < 2^(t+1)  ->  pre = t + 1, so that M = floor(2^(32-pre) g/d) + 1 lies in [2^31, 2^32)
        unsigned t = 0;
        while (((uint64_t)d << (t + 1)) <= (uint64_t)g)
            t++;
        pre = t + 1;
    }
    const unsigned k = 32 - pre;
    const uint64_t M = (((uint64_t)g << k) / d) + 1;       // g < 2^16, k <= 32: fits
    RecipeHost r;
    r.mw = (uint32_t)M;
    r.addm = (M >> 31) ? 0xffffffffu : 0u;                  // M < 2^32 always
    r.mul = 1u << pre;
    return r;
}

int recipe_eval(const RecipeHost &r, int x)
{
    // bit-for-bit what apply_gain<GM_MASKED>() does on the device, in plain integer C
    const int32_t X = (int32_t)((uint32_t)x * r.mul);
    const uint64_t add = ((uint64_t)((uint32_t)X & r.addm) << 32) | (X < 0 ? 0xffffffffull : 0ull);
    const uint64_t acc = (uint64_t)((int64_t)X * (int64_t)(int32_t)r.mw) + add;       // wraps like the device's 64-bit add
    int32_t y = (int32_t)(uint32_t)(acc >> 32);
    if (y > 32767)
        y = 32767;
    if (y < -32768)
        y = -32768;
    return y;
}

unsigned ceil_log2(uint32_t v)
{
    unsigned b = 0;
    while ((1ull << b) < v)
        b++;
    return b;
}

}  // namespace

namespace {

using TickKernel = void (*)(const TickArgs);

template <int C, int G>
TickKernel fast_kernel(int gm, bool meter, bool planar, bool nc)
{
    using namespace cmgpu;
    if (nc && !planar) {    // separate output ring: loads through the read-only path
        switch (gm) {
        case GM_IDENTITY: return meter ? fused_tick<C, G, GM_IDENTITY, true, false, true> : fused_tick<C, G, GM_IDENTITY, false, false, true>;
        case GM_ADDALL:   return meter ? fused_tick<C, G, GM_ADDALL, true, false, true> : fused_tick<C, G, GM_ADDALL, false, false, true>;
        default:          return meter ? fused_tick<C, G, GM_MASKED, true, false, true> : fused_tick<C, G, GM_MASKED, false, false, true>;
        }
    }
    if (planar) {           // planes are a second output of the fused (metering) pass only
        switch (gm) {
        case GM_IDENTITY: return fused_tick<C, G, GM_IDENTITY, true, true>;
        case GM_ADDALL:   return fused_tick<C, G, GM_ADDALL, true, true>;
        default:          return fused_tick<C, G, GM_MASKED, true, true>;
        }
    }
    switch (gm) {
    case GM_IDENTITY: return meter ? fused_tick<C, G, GM_IDENTITY, true> : fused_tick<C, G, GM_IDENTITY, false>;
    case GM_ADDALL:   return meter ? fused_tick<C, G, GM_ADDALL, true> : fused_tick<C, G, GM_ADDALL, false>;
    default:          return meter ? fused_tick<C, G, GM_MASKED, true> : fused_tick<C, G, GM_MASKED, false>;
    }
}

template <int C>
TickKernel fast_kernel_g(int g, int gm, bool meter, bool planar, bool nc)
{
    if (g == 8 && C != 16)
        return fast_kernel<C, (C == 16 ? 32 : 8)>(gm, meter, planar, nc);
    return fast_kernel<C, 32>(gm, meter, planar, nc);
}

using GenericKernel = void (*)(const TickArgs, const int);

GenericKernel generic_kernel(int gm, bool meter)
{
    using namespace cmgpu;
    switch (gm) {
    case GM_IDENTITY: return meter ? generic_tick<GM_IDENTITY, true> : generic_tick<GM_IDENTITY, false>;
    // the generic kernel has no add-all specialisation: the masked recipe covers it
    default:          return meter ? generic_tick<GM_MASKED, true> : generic_tick<GM_MASKED, false>;
    }
}

template <int C>
TickKernel tma_kernel_c(int gm, bool meter)
{
    using namespace cmgpu;
    switch (gm) {
    case GM_IDENTITY: return meter ? tma_tick<C, GM_IDENTITY, true> : tma_tick<C, GM_IDENTITY, false>;
    case GM_ADDALL:   return meter ? tma_tick<C, GM_ADDALL, true> : tma_tick<C, GM_ADDALL, false>;
    default:          return meter ? tma_tick<C, GM_MASKED, true> : tma_tick<C, GM_MASKED, false>;
    }
}
TickKernel tma_kernel(const cmgpu_ctx *c, int gm, bool meter)
{
    switch (c->channels) {
    case 1:  return tma_kernel_c<1>(gm, meter);
    case 2:  return tma_kernel_c<2>(gm, meter);
    case 4:  return tma_kernel_c<4>(gm, meter);
    case 8:  return tma_kernel_c<8>(gm, meter);
    default: return tma_kernel_c<16>(gm, meter);
    }
}

using SpanKernel = void (*)(const TickArgs, const uint32_t);

template <int C>
SpanKernel span_kernel_c(int gm, bool meter, bool nc)
{
    using namespace cmgpu;
    if (nc) {
        switch (gm) {
        case GM_IDENTITY: return meter ? span_tick<C, GM_IDENTITY, true, true> : span_tick<C, GM_IDENTITY, false, true>;
        case GM_ADDALL:   return meter ? span_tick<C, GM_ADDALL, true, true> : span_tick<C, GM_ADDALL, false, true>;
        default:          return meter ? span_tick<C, GM_MASKED, true, true> : span_tick<C, GM_MASKED, false, true>;
        }
    }
    switch (gm) {
    case GM_IDENTITY: return meter ? span_tick<C, GM_IDENTITY, true, false> : span_tick<C, GM_IDENTITY, false, false>;
    case GM_ADDALL:   return meter ? span_tick<C, GM_ADDALL, true, false> : span_tick<C, GM_ADDALL, false, false>;
    default:          return meter ? span_tick<C, GM_MASKED, true, false> : span_tick<C, GM_MASKED, false, false>;
    }
}
SpanKernel span_kernel(const cmgpu_ctx *c, int gm, bool meter)
{
    const bool nc = c->d_out != nullptr;
    switch (c->channels) {
    case 1:  return span_kernel_c<1>(gm, meter, nc);
    case 2:  return span_kernel_c<2>(gm, meter, nc);
    case 4:  return span_kernel_c<4>(gm, meter, nc);
    default: return span_kernel_c<8>(gm, meter, nc);
    }
}

using AnyKernel = void (*)(const TickArgs, const int, const int);

AnyKernel any_kernel(int gm, bool meter, bool nc)
{
    using namespace cmgpu;
    if (nc) {               // separate output ring: loads through the read-only path
        switch (gm) {
        case GM_IDENTITY: return meter ? any_tick<GM_IDENTITY, true, true> : any_tick<GM_IDENTITY, false, true>;
        case GM_ADDALL:   return meter ? any_tick<GM_ADDALL, true, true> : any_tick<GM_ADDALL, false, true>;
        default:          return meter ? any_tick<GM_MASKED, true, true> : any_tick<GM_MASKED, false, true>;
        }
    }
    switch (gm) {
    case GM_IDENTITY: return meter ? any_tick<GM_IDENTITY, true> : any_tick<GM_IDENTITY, false>;
    case GM_ADDALL:   return meter ? any_tick<GM_ADDALL, true> : any_tick<GM_ADDALL, false>;
    default:          return meter ? any_tick<GM_MASKED, true> : any_tick<GM_MASKED, false>;
    }
}

TickKernel pick_fast(const cmgpu_ctx *c, int gm, bool meter, bool planar = false)
{
    const bool nc = c->d_out != nullptr;        // CMGPU_SEPARATE_OUT: the input ring is read-only for a tick
    switch (c->channels) {
    case 1:  return fast_kernel_g<1>(c->plan_g, gm, meter, planar, nc);
    case 2:  return fast_kernel_g<2>(c->plan_g, gm, meter, planar, nc);
    case 4:  return fast_kernel_g<4>(c->plan_g, gm, meter, planar, nc);
    case 8:  return fast_kernel_g<8>(c->plan_g, gm, meter, planar, nc);
    default: return fast_kernel_g<16>(c->plan_g, gm, meter, planar, nc);
    }
}

int resident_ctas(cmgpu_ctx *c, int gm, bool meter)
{
    int &cap = c->grid_cap[gm][meter ? 1 : 0];
    if (cap)
        return cap;
    int n = 0;
    cudaError_t e = c->plan_g > 0   ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, pick_fast(c, gm, meter), 256, 0)
                    : c->plan_g < 0 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, any_kernel(gm, meter, c->d_out != nullptr), 256,
                                                                                    cmgpu::kAnySmemBytes)
                                    : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, generic_kernel(gm, meter), 128, 0);
    if (e != cudaSuccess || n < 1)
        n = 1;
    cap = n * c->num_sms;
    return cap;
}

// `pdl`: the launch may start while the previous launch of the stream is still draining
// (programmatic dependent launch; the kernels' launch_begin / launch_end are the device side).
template <typename... KArgs, typename... Args>
cudaError_t launch_kernel(void (*k)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t st, bool pdl,
                          Args... args)
{
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, k, KArgs(args)...);
}

cudaError_t launch_tick(cmgpu_ctx *c, const TickArgs &a, int gm, bool meter, cudaStream_t st, bool pdl)
{
    if (c->tma && !a.planar) {
        TickKernel k = tma_kernel(c, gm, meter);
        int &cap = c->tma_grid_cap[gm][meter ? 1 : 0];
        if (!cap) {
            cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, cmgpu::kTmaSmemBytes);
            if (e != cudaSuccess)
                return e;
            int n = 0;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k, cmgpu::kTmaThreads, cmgpu::kTmaSmemBytes) != cudaSuccess || n < 1)
                n = 1;
            cap = n * c->num_sms;
        }
        TickArgs t = a;
        t.items_per_block = c->tma_items;
        t.per_item = c->tma_per_item;
        uint64_t grid = (uint64_t)t.n_streams * t.items_per_block;
        if (grid > (uint64_t)cap)
            grid = (uint64_t)cap;
        return launch_kernel(k, (unsigned)grid, cmgpu::kTmaThreads, cmgpu::kTmaSmemBytes, st, pdl, t);
    }
    // Small-buffer spans with enough streams to fill the machine with 8-lane groups: one group per STREAM
    // walks all ticks of the span and publishes its meter partials once (cmgpu_span.cuh)
    if ((a.n_ticks > 1 || c->env_span_single) && c->plan_g == 8 && c->channels <= 8 && !a.planar && a.n_ticks <= cmgpu::kSpanMaxTicks &&
        !c->env_span_by_tick && (c->env_span_by_stream || (uint64_t)a.n_streams * 8u * 2u >= (uint64_t)c->num_sms * 1024u)) {
        SpanKernel k = span_kernel(c, gm, meter);
        const uint32_t vmax = (uint32_t)((c->stride / 16 + 7) / 8);          // vectors of a stream-block per lane, <= 8
        const size_t smem = (size_t)2 * vmax * 256 * 16;
        int &cap = c->span_grid_cap[gm][meter ? 1 : 0];
        if (!cap) {
            cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess)
                return e;
            int n = 0;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k, 256, smem) != cudaSuccess || n < 1)
                n = 1;
            cap = n * c->num_sms;
        }
        uint64_t grid = ((uint64_t)a.n_streams + 31u) / 32u;
        if (grid > (uint64_t)cap)
            grid = (uint64_t)cap;
        snprintf(c->kname_last, sizeof(c->kname_last), "span_tick<C=%u> (%u ticks per launch)", c->channels, a.n_ticks);
        return launch_kernel(k, (unsigned)grid, 256, smem, st, pdl, a, vmax);
    }
    c->kname_last[0] = 0;
    const uint64_t items = (uint64_t)a.n_streams * a.items_per_block * (a.n_ticks > 1 ? a.n_ticks : 1u);
    const uint64_t per_cta = c->plan_g > 0 ? 256u / (unsigned)c->plan_g : (c->plan_g < 0 ? 8u : 4u);
    uint64_t grid = (items + per_cta - 1) / per_cta;
    const uint64_t cap = (uint64_t)resident_ctas(c, gm, meter);
    if (grid > cap)
        grid = cap;
    if (c->plan_g < 0)
        return launch_kernel(any_kernel(gm, meter, c->d_out != nullptr), (unsigned)grid, 256, cmgpu::kAnySmemBytes, st, pdl, a,
                             (int)c->channels, c->plan_lanes);
    if (c->plan_g == 0)
        return launch_kernel(generic_kernel(gm, meter), (unsigned)grid, 128, 0, st, pdl, a, (int)c->channels);
    return launch_kernel(pick_fast(c, gm, meter, a.planar != nullptr), (unsigned)grid, 256, 0, st, pdl, a);
}

// Decide how a tick is cut into work items. Shape-only, so done once per context.
void make_plan(cmgpu_ctx *c)
{
    const unsigned C = c->channels;
    const bool fast = !(c->flags & CMGPU_FORCE_GENERIC) && (C == 1 || C == 2 || C == 4 || C == 8 || C == 16);
    if (!fast && !(c->flags & CMGPU_FORCE_GENERIC)) {
        // any_tick: vectors over frames that straddle them; L lanes, 8*L a multiple of C
        unsigned g = 8, x = C;
        while (x) { unsigned t = g % x; g = x; x = t; }          // gcd(8, C)
        const unsigned m = C / g;
        const unsigned lanes = 32 - 32 % m;
        const uint32_t nvec = (uint32_t)(c->stride / 16);
        const uint32_t quantum = lanes * 8u;                      // two batches of 4 per lane
        // Large items: every item boundary costs an epilogue and a recipe gather. Measured on 6-channel
        // streams with the cp.async ring (profiles/r2_*cfg6ch*): 2,048 vectors per item 0.89 of
        // the copy rate, 4,096 0.94, 7,000 0.985; the epilogue's 16-bit sample index caps an item at 8,192.
        uint32_t target = 7000;
        if (const char *e = getenv("CMGPU_ITEM_VECS"))            // tuning hook
            target = (uint32_t)strtoul(e, nullptr, 10) ? (uint32_t)strtoul(e, nullptr, 10) : target;
        if (target > 8192)                                        // the epilogue's sample index (vector * 8 + slot) fits 16 bits
            target = 8192 - quantum;
        uint32_t items = (nvec + target - 1) / target;
        uint32_t per = (nvec + items - 1) / items;
        per = (per + quantum - 1) / quantum * quantum;
        items = (nvec + per - 1) / per;
        c->plan_g = -1;
        c->plan_lanes = (int)lanes;
        c->plan_items = items;
        c->plan_per_item = per;
        snprintf(c->kname, sizeof(c->kname), "any_tick<C=%u,L=%u>", C, lanes);
        return;
    }
    if (!fast) {
        // generic: one warp per item, 32 frames per step; aim for <= 2048 frames per item
        const uint32_t target = 2048;
        uint32_t items = (c->block_frames + target - 1) / target;
        uint32_t per = (c->block_frames + items - 1) / items;
        per = (per + 31u) & ~31u;
        items = (c->block_frames + per - 1) / per;
        c->plan_g = 0;
        c->plan_items = items;
        c->plan_per_item = per;
        snprintf(c->kname, sizeof(c->kname), "generic_tick<C=%u>", C);
        return;
    }
    const uint32_t nvec = (uint32_t)(c->stride / 16);
    // stream-blocks of up to 1 KiB are walked by 8-lane groups so that no lane idles
    const int g = (nvec <= 64 && C != 16) ? 8 : 32;
    // Aim for 24 KiB (1,536 vectors) per item for mono / stereo and 32 KiB (2,048) for wider frames, a
    // multiple of 4 steps of the group. Measured, not derived (tools/sweep_item.sh on two boxes, with
    // overlapping launches): stereo cfg2 0.620 ms at 2,048, 0.598-0.601 at 1,536, 0.607 at 1,280, 0.604 at
    // 4,096; cfg5 4.11 / 4.01 / 4.04 ms at 2,048 / 1,536 / 1,280; 8-channel cfg4a 1.98 at 1,536, 1.93 at
    // 2,048, 1.92 at 4,096.
    uint32_t target = C <= 2 ? 1536 : 2048;
    if (const char *e = getenv("CMGPU_ITEM_VECS"))          // tuning hook
        target = (uint32_t)strtoul(e, nullptr, 10) ? (uint32_t)strtoul(e, nullptr, 10) : target;
    if (target > 32768)                                       // item_publish: vector indices of an item fit 16 bits
        target = 32768;
    uint32_t items = (nvec + target - 1) / target;
    uint32_t per = (nvec + items - 1) / items;
    const uint32_t quantum = (uint32_t)g * 4u;
    per = (per + quantum - 1) / quantum * quantum;
    items = (nvec + per - 1) / per;
    c->plan_g = g;
    c->plan_items = items;
    c->plan_per_item = per;
    // Opt-in (CMGPU_TMA=1): stream-blocks of >= 128 KiB with their loads staged through shared memory
    // by TMA bulk copies (cmgpu_tma.cuh). Measured on B200: as a pure copy it equals the
    // LDG kernel (0.650 ms on cfg2), with gain + meter it is slower (0.71 vs 0.62-0.65), so the LDG
    // kernel stays the default.
    if (nvec >= 8192 && getenv("CMGPU_TMA")) {
        uint32_t tile_target = 8 * cmgpu::kTmaTileVecs;
        if (const char *e = getenv("CMGPU_TMA_ITEM_TILES"))         // tuning hook
            tile_target = (uint32_t)(strtoul(e, nullptr, 10) ? strtoul(e, nullptr, 10) : 8) * cmgpu::kTmaTileVecs;
        if (tile_target > 32768)                                    // item_publish: vector indices of an item fit 16 bits
            tile_target = 32768;
        uint32_t n = (nvec + tile_target - 1) / tile_target;
        uint32_t tper = (nvec + n - 1) / n;
        tper = (tper + cmgpu::kTmaTileVecs - 1) / cmgpu::kTmaTileVecs * cmgpu::kTmaTileVecs;
        c->tma = true;
        c->tma_items = (nvec + tper - 1) / tper;
        c->tma_per_item = tper;
        snprintf(c->kname, sizeof(c->kname), "tma_tick<C=%u>", C);
        return;
    }
    snprintf(c->kname, sizeof(c->kname), "fused_tick<C=%u,G=%d>", C, g);
}

int upload_gains_locked(cmgpu_ctx *c)
{
    if (c->dirty_lo >= c->dirty_hi)
        return CMGPU_OK;
    // pageable source: the runtime stages it before returning, later edits cannot race
    CU(cudaMemcpyAsync(c->d_gains + c->dirty_lo, c->h_gains.data() + c->dirty_lo,
                       sizeof(GainRow) * (c->dirty_hi - c->dirty_lo), cudaMemcpyHostToDevice, c->cmp()));
    c->dirty_lo = c->max_streams;
    c->dirty_hi = 0;
    return CMGPU_OK;
}

void mark_dirty(cmgpu_ctx *c, unsigned lo, unsigned hi)
{
    c->classes_dirty = true;
    c->config_gen++;
    if (lo < c->dirty_lo)
        c->dirty_lo = lo;
    if (hi > c->dirty_hi)
        c->dirty_hi = hi;
}

void set_row(cmgpu_ctx *c, unsigned s, uint16_t scale, const uint16_t *gain)
{
    GainRow &r = c->h_gains[s];
    memset(&r, 0, sizeof(r));
    c->h_scale[s] = scale;
    bool identity = (scale == 0);
    if (scale) {
        bool unity = true;
        for (unsigned ch = 0; ch < c->channels; ch++) {
            c->h_gain[(size_t)s * c->channels + ch] = gain[ch];
            RecipeHost h = make_recipe(gain[ch], scale);
            r.mw[ch] = h.mw;
            r.addm[ch] = h.addm;
            r.mul[ch] = h.mul;
            unity = unity && gain[ch] == scale;
        }
        identity = unity;     // trunc(x*d/d) == x and x is already inside the clamp range
    }
    // rows the identity kernel skips still carry an exact recipe (g == d == 1), so that they can
    // ride along in a gain-mode launch when other streams of the tick need one
    const RecipeHost unity = make_recipe(1, 1);
    for (unsigned ch = scale ? c->channels : 0; ch < 16; ch++) {
        r.mw[ch] = unity.mw;
        r.addm[ch] = unity.addm;
        r.mul[ch] = unity.mul;
    }
    bool addall = scale != 0;
    for (unsigned ch = 0; ch < c->channels && addall; ch++)
        addall = r.addm[ch] == 0xffffffffu;
    r.flags = (identity ? cmgpu::kGainIdentity : 0) | (addall ? cmgpu::kGainAddAll : 0);
}

int gain_mode_of(const GainRow &r)
{
    if (r.flags & cmgpu::kGainIdentity)
        return cmgpu::GM_IDENTITY;
    return (r.flags & cmgpu::kGainAddAll) ? cmgpu::GM_ADDALL : cmgpu::GM_MASKED;
}

int rebuild_classes_locked(cmgpu_ctx *c)
{
    c->n_mode[0] = c->n_mode[1] = c->n_mode[2] = 0;
    for (unsigned s = 0; s < c->active; s++)
        c->n_mode[gain_mode_of(c->h_gains[s])]++;
    c->classes_dirty = false;
    return CMGPU_OK;
}

int flush_ticks_locked(cmgpu_ctx *c);

// Dynamic work claims (TickArgs::work) for a launch of `n_items` warp-sized items on a grid of `grid`
// CTAs out of `cap` resident ones. Every launch takes the next of kWorkCounters counters, so launches
// that overlap never share one: with a grid of at least a quarter of the resident CTAs no more than five
// launches fit on the GPU at once (a launch cannot finish before its predecessor has). Smaller
// overlapping grids, spans, captured cycles (their ticks run side by side on forked streams and are
// replayed with the same arguments) stay static.
void assign_work(cmgpu_ctx *c, bool pdl, bool captured, cudaStream_t st, uint64_t n_items, uint64_t grid, uint64_t cap,
                 unsigned int **work, uint32_t *base)
{
    *work = nullptr;
    *base = 0;
    if (captured || st != c->s_cmp || c->env_static || n_items >= 0xffffffffull)
        return;
    if (pdl && grid * 4u < cap)
        return;
    const unsigned k = c->work_next++ % cmgpu_ctx::kWorkCounters;
    *work = c->d_work + k;
    *base = c->work_base[k];
    c->work_base[k] += (uint32_t)n_items;              // one claim per processed item (mod 2^32, like the counter)
}

#ifdef CMGPU_BOUNDS_CHECK
int debug_set_bounds(cmgpu_ctx *c, cudaStream_t st)
{
    cmgpu::DebugBounds b;
    const size_t ring = c->slot_bytes * c->slots;
    const size_t ring_out = c->out_channels ? c->slot_bytes_out * c->slots : ring;
    b.lo[0] = (unsigned long long)c->d_in;
    b.hi[0] = b.lo[0] + ring;
    b.lo[1] = (unsigned long long)(c->d_out ? c->d_out : c->d_in);
    b.hi[1] = b.lo[1] + (c->d_out ? ring_out : ring);
    // pageable source: staged by the runtime before the call returns
    CU(cudaMemcpyToSymbolAsync(cmgpu::g_dbg_bounds, &b, sizeof(b), 0, cudaMemcpyHostToDevice, st));
    return CMGPU_OK;
}
#endif

// One tick on stream `st`. In a cycle (cmgpu_process_cycle) the ticks run concurrently: each gets
// its place in the sequence as `tick_offset` and leaves advancing the counter to the cycle's end.
// n_ticks > 1: a span -- ONE launch over the consecutive slots [slot, slot + n_ticks) (span_ok() says when).
// `captured`: the launch is being recorded into a cycle's graph; its tick number is then relative to
// the device counter (tick_offset = place in the cycle) and the graph's own bump_tick advances it.
// Otherwise the host numbers the tick(s): offset = ticks issued since the last bump.
int launch_locked(cmgpu_ctx *c, unsigned slot, unsigned flags, cudaStream_t st = nullptr, unsigned tick_offset = 0,
                  bool captured = false, unsigned n_ticks = 1)
{
    if (!st)
        st = c->cmp();
    int rc = upload_gains_locked(c);
    if (rc)
        return rc;
    if (!captured) {
        // the host numbers ticks in 32 bits on top of the device's 64-bit counter: fold them in long
        // before the offset can wrap (a full dependency once every 2^31 ticks)
        if (c->pending_ticks >= 0x80000000u && (rc = flush_ticks_locked(c)))
            return rc;
        tick_offset = c->pending_ticks;
    }
    const bool meter = (flags & CMGPU_METER) != 0;
    const bool transform = (flags & CMGPU_TRANSFORM) != 0;
    if (!c->active)
        return CMGPU_OK;
    if (c->out_channels) {
        // EXTENSION: downmix contexts always mix and meter both sides
        if (c->mix_dirty) {
            CU(cudaMemcpyAsync(c->d_mix, c->h_mix.data(), sizeof(MixRow) * c->max_streams, cudaMemcpyHostToDevice, st));
            c->mix_dirty = false;
        }
        MixArgs m;
        memset(&m, 0, sizeof(m));
        m.in = c->d_in + (size_t)slot * c->slot_bytes;
        m.out = c->d_out + (size_t)slot * c->slot_bytes_out;
        m.frames = c->has_frames[slot] ? c->d_frames + (size_t)slot * c->max_streams : nullptr;
        m.rows = c->d_mix;
        m.meters_in = c->d_meters_in;
        m.meters_out = c->d_meters;
        m.tick = c->d_tick;
        m.pbits = c->pbits;
        m.tick_offset = tick_offset;
        m.n_streams = c->active;
        m.block_frames = c->block_frames;
        m.stride_in = (uint32_t)c->stride;
        m.stride_out = (uint32_t)c->stride_out;
        const uint32_t target = 2048;
        uint32_t items = (c->block_frames + target - 1) / target;
        uint32_t per = ((c->block_frames + items - 1) / items + 255u) & ~255u;
        m.items_per_block = (c->block_frames + per - 1) / per;
        m.per_item = per;
        m.cin = c->channels;
        m.cout = c->out_channels;
        const bool vec8 = c->channels == 8 && c->out_channels == 2 && !(c->flags & CMGPU_FORCE_GENERIC);
        const uint64_t n_items = (uint64_t)m.n_streams * m.items_per_block;
        const unsigned per_cta = vec8 ? 8 : 4;
        uint64_t grid = (n_items + per_cta - 1) / per_cta;
        int occ = 0;
        const bool in_meter = !(c->flags & CMGPU_MIX_OUTPUT_METER_ONLY);
        if ((vec8 ? (in_meter ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, cmgpu::mix8to2_tick<true>, 256, 0)
                              : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, cmgpu::mix8to2_tick<false>, 256, 0))
                  : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, cmgpu::mix_tick<false>, 128, 0)) != cudaSuccess ||
            occ < 1)
            occ = 1;
        if (grid > (uint64_t)occ * c->num_sms)
            grid = (uint64_t)occ * c->num_sms;
        // a downmix context always has its own output ring: consecutive ticks never conflict
        const bool mpdl = !captured && st == c->s_cmp && c->chain_open && c->last_first != ~0u && !c->env_no_pdl;
        if (!captured && st == c->s_cmp)
            c->chain_open = true;
#ifdef CMGPU_BOUNDS_CHECK
        if (!captured)
            if (int brc = debug_set_bounds(c, st))
                return brc;
#endif
        assign_work(c, mpdl, captured, st, n_items, grid, (uint64_t)occ * c->num_sms, &m.work, &m.work_base);
        // completion word, as for the plain ticks below
        const bool mflagged = !captured && st == c->s_cmp && c->h_done != nullptr && c->idle_hint.exchange(false, std::memory_order_acq_rel);
        if (mflagged) {
            m.done_count = c->d_done_count;
            m.done_flag = c->d_done_flag;
            m.done_gen = c->done_gen_next;
        }
        if (vec8 && in_meter)
            CU(launch_kernel(cmgpu::mix8to2_tick<true>, (unsigned)grid, 256, 0, st, mpdl, m));
        else if (vec8)
            CU(launch_kernel(cmgpu::mix8to2_tick<false>, (unsigned)grid, 256, 0, st, mpdl, m));
        else
            CU(launch_kernel(cmgpu::mix_tick<false>, (unsigned)grid, 128, 0, st, mpdl, m));
        c->launches++;
        if (!captured) {
            c->pending_ticks += 1;
            if (st == c->s_cmp)
                c->last_first = slot;
        }
        if (mflagged) {
            c->tail_gen.store(m.done_gen, std::memory_order_release);
            if (++c->done_gen_next == 0)
                c->done_gen_next = 1;
        }
        return CMGPU_OK;
    }
    if (c->classes_dirty && (rc = rebuild_classes_locked(c)))
        return rc;
    const bool separate = c->d_out != nullptr;
    // the cheapest mode that is exact for every active stream
    int gm = cmgpu::GM_IDENTITY;
    if (transform && c->n_mode[cmgpu::GM_MASKED])
        gm = cmgpu::GM_MASKED;
    else if (transform && c->n_mode[cmgpu::GM_ADDALL])
        gm = (c->plan_g == 0) ? cmgpu::GM_MASKED : cmgpu::GM_ADDALL;   // generic kernel: masked covers it
    const bool store = gm != cmgpu::GM_IDENTITY || separate;
    const bool planar = (flags & CMGPU_PLANAR) != 0;
    if (planar && (!c->d_planar || !meter))
        return fail(CMGPU_ERR_INVAL, "CMGPU_PLANAR needs a CMGPU_PLANAR_F32 context and goes with CMGPU_METER (the fused pass)");
    if (!store && !meter && !planar)
        return CMGPU_OK;                 // in-place pass-through without metering: nothing to do

    TickArgs a;
    memset(&a, 0, sizeof(a));
    a.in = c->d_in + (size_t)slot * c->slot_bytes;
    a.out = (c->d_out ? c->d_out : c->d_in) + (size_t)slot * c->slot_bytes;
    a.frames = c->has_frames[slot] ? c->d_frames + (size_t)slot * c->max_streams : nullptr;
    a.gains = c->d_gains;
    a.meters = c->d_meters;
    a.tick = c->d_tick;
    a.pbits = c->pbits;
    a.tick_offset = tick_offset;
    a.n_streams = c->active;
    a.block_frames = c->block_frames;
    a.stride_bytes = (uint32_t)c->stride;
    a.items_per_block = c->plan_items;
    a.per_item = c->plan_per_item;
    a.row_u64 = c->row_u64;
    a.store = store ? 1u : 0u;
    a.planar = planar ? c->d_planar + (size_t)slot * c->planar_slot_floats : nullptr;
    a.plane_stride = (uint32_t)c->plane_stride;
    if (store && !separate)
        for (unsigned i = slot; i < slot + n_ticks && i < c->slots; i++)
            c->slot_dirty[i] = 1;                    // the device copy now differs from what was uploaded
    a.n_ticks = n_ticks;
    a.frames_stride = c->max_streams;
    a.slot_bytes = c->slot_bytes;
    // A launch may start while earlier tick launches of the compute stream are still running when it
    // reads nothing any of them writes: slots none of the still-open chain touches, or a separate
    // output ring (the input ring is then read-only for ticks). The host's model is never less strict
    // than the hardware: in_chain[] is cleared ONLY by a launch issued WITHOUT the overlap attribute
    // (which the stream orders after everything before it), so whatever else was queued in between
    // (uploads of gains or frame counts, snapshots, fills, event waits -- each an ordinary full
    // dependency on the device) can only make the real overlap smaller than the assumed one. Entry
    // points that queue a non-tick kernel also clear chain_open, so the next tick carries no attribute.
    bool pdl = false;
    if (!captured && st == c->s_cmp) {
        bool conflict = false;
        if (!separate)
            for (unsigned i = slot; i < slot + n_ticks && i < c->slots; i++)
                conflict = conflict || c->in_chain[i];
        pdl = c->chain_open && c->last_first != ~0u && !conflict && !c->env_no_pdl;
        if (!pdl)
            std::fill(c->in_chain.begin(), c->in_chain.end(), 0);      // a full dependency: everything before is done
        for (unsigned i = slot; i < slot + n_ticks && i < c->slots; i++)
            c->in_chain[i] = 1;
        c->chain_open = true;
    }
    // plain ticks of the warp-per-item kernels claim their work items dynamically (assign_work)
    a.work = nullptr;
    a.work_base = 0;
    if (n_ticks <= 1 && c->plan_g != 8 && !c->tma) {
        const uint64_t n_items = (uint64_t)a.n_streams * a.items_per_block;
        const uint64_t per_cta = c->plan_g == 0 ? 4u : 8u;
        const uint64_t cap = (uint64_t)resident_ctas(c, gm, meter);
        const uint64_t grid = std::min<uint64_t>((n_items + per_cta - 1) / per_cta, cap);
        assign_work(c, pdl, captured, st, n_items, grid, cap, &a.work, &a.work_base);
    }
#ifdef CMGPU_BOUNDS_CHECK
    if (!captured)                      // (a captured copy from host memory would be replayed from a dead stack frame)
        if (int brc = debug_set_bounds(c, st))
            return brc;
#endif
    // completion word: the tail of the compute stream is now this launch (cmgpu_sync polls the word)
    // Only a launch queued on a compute stream somebody has just waited for gets one: the word costs the
    // launch an atomic per CTA and a write across the host link at its very end (3 us on a 10 us tick,
    // serialised from launch to launch in a back-to-back burst: 8.2 -> 11.1 us per config-3 tick), and it
    // only pays where the host waits tick by tick (17.5 -> 14.4 us launch-to-complete there).
    const bool flagged = !captured && st == c->s_cmp && c->h_done != nullptr && c->idle_hint.exchange(false, std::memory_order_acq_rel);
    if (flagged) {
        a.done_count = c->d_done_count;
        a.done_flag = c->d_done_flag;
        a.done_gen = c->done_gen_next;
    }
    CU(launch_tick(c, a, gm, meter, st, pdl));
    c->launches++;
    if (!captured) {
        c->pending_ticks += n_ticks;
        if (st == c->s_cmp)
            c->last_first = slot;
    }
    if (flagged) {
        c->tail_gen.store(a.done_gen, std::memory_order_release);
        if (++c->done_gen_next == 0)
            c->done_gen_next = 1;
    }
    return CMGPU_OK;
}

bool slot_ok(const cmgpu_ctx *c, unsigned slot) { return c && slot < c->slots; }

// "every tick queued on this slot so far": recorded on demand (a superset -- everything queued on the
// compute stream so far -- which is what the next upload / download of the slot has to wait for).
int ticks_done_event_locked(cmgpu_ctx *c, unsigned slot)
{
    if (c->cmp_unrecorded[slot]) {
        CU(cudaEventRecord(c->ev_cmp[slot], c->cmp()));
        c->cmp_unrecorded[slot] = 0;
        c->last_first = ~0u;
        c->chain_open = false;
    }
    return CMGPU_OK;
}

int tick_waits_locked(cmgpu_ctx *c, unsigned slot)
{
    if (c->up_pending[slot]) {
        CU(cudaStreamWaitEvent(c->cmp(), c->ev_up[slot], 0));
        c->up_pending[slot] = 0;
    }
    if (c->down_pending[slot]) {
        CU(cudaStreamWaitEvent(c->cmp(), c->ev_down[slot], 0));
        c->down_pending[slot] = 0;
    }
    return CMGPU_OK;
}

}  // namespace

extern "C" {

const char *cmgpu_version(void) { return "coolmic-b200 0.2 (sm_100a)"; }

const char *cmgpu_last_error(void) { return cmgpu::g_err; }

int cmgpu_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

void *cmgpu_host_alloc(size_t bytes)
{
    void *p = nullptr;
    cudaError_t e = cudaMallocHost(&p, bytes ? bytes : 1);
    if (e != cudaSuccess) {
        fail(CMGPU_ERR_NOMEM, "cudaMallocHost(%zu): %s", bytes, cudaGetErrorString(e));
        return nullptr;
    }
    return p;
}

void *cmgpu_host_alloc_wc(size_t bytes)
{
    void *p = nullptr;
    cudaError_t e = cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocWriteCombined);
    if (e != cudaSuccess) {
        fail(CMGPU_ERR_NOMEM, "cudaHostAlloc(%zu, write-combined): %s", bytes, cudaGetErrorString(e));
        return nullptr;
    }
    return p;
}

void cmgpu_host_free(void *p)
{
    if (p)
        cudaFreeHost(p);
}

static cmgpu_ctx_t *ctx_create_impl(int device, unsigned channels, unsigned out_channels, unsigned max_streams,
                                    unsigned ring_slots, unsigned block_frames, unsigned flags)
{
    if (!channels || channels > CMGPU_MAX_CHANNELS || out_channels > CMGPU_MAX_CHANNELS || !max_streams || !ring_slots ||
        !block_frames) {
        fail(CMGPU_ERR_INVAL, "cmgpu_ctx_create: channels 1..16, streams, slots and block_frames must be non-zero");
        return nullptr;
    }
    const uint64_t raw = (uint64_t)block_frames * channels * 2u;
    if (raw > 0x7ffffff0ull) {
        fail(CMGPU_ERR_INVAL, "cmgpu_ctx_create: stream-block of %llu bytes is too large", (unsigned long long)raw);
        return nullptr;
    }
    int ndev = cmgpu_device_count();
    if (device < 0 || device >= ndev) {
        fail(CMGPU_ERR_GENERIC, "cmgpu_ctx_create: CUDA device %d not available (%d devices); there is no CPU fallback",
             device, ndev);
        return nullptr;
    }
    cmgpu_ctx *c = new (std::nothrow) cmgpu_ctx;
    if (!c) {
        fail(CMGPU_ERR_NOMEM, "out of host memory");
        return nullptr;
    }
    c->device = device;
    c->env_no_pdl = getenv("CMGPU_NO_PDL") != nullptr;
    c->env_no_span = getenv("CMGPU_NO_SPAN") != nullptr;
    c->env_static = getenv("CMGPU_STATIC_ITEMS") != nullptr;
    c->env_span_by_tick = getenv("CMGPU_SPAN_BY_TICK") != nullptr;    // A/B hook: spans as (tick, stream) work items
    c->env_span_by_stream = getenv("CMGPU_SPAN_BY_STREAM") != nullptr;
    c->env_span_single = getenv("CMGPU_SPAN_SINGLE") != nullptr;        // A/B hook: single ticks through the span kernel too    // test hook: stream-major spans for any stream count
    c->channels = channels;
    c->max_streams = c->active = max_streams;
    c->slots = ring_slots;
    c->block_frames = block_frames;
    c->flags = flags;
    c->stride = (size_t)((raw + 15u) & ~15ull);
    c->slot_bytes = c->stride * max_streams;
    c->row_u64 = 2 * channels + 2;
    c->pbits = ceil_log2(block_frames) ? ceil_log2(block_frames) : 1;
    if (out_channels) {
        // downmix context: the meter table a caller sees is the OUTPUT side's
        c->out_channels = out_channels;
        c->stride_out = (size_t)(((uint64_t)block_frames * out_channels * 2u + 15u) & ~15ull);
        c->slot_bytes_out = c->stride_out * max_streams;
        c->row_in_u64 = 2 * channels + 2;
        c->row_u64 = 2 * out_channels + 2;
        flags &= ~CMGPU_SEPARATE_OUT;
        c->flags = flags;
    }

    auto bail = [&](const char *what, cudaError_t e) -> cmgpu_ctx_t * {
        fail(e == cudaErrorMemoryAllocation ? CMGPU_ERR_NOMEM : CMGPU_ERR_GENERIC, "cmgpu_ctx_create: %s: %s", what,
             cudaGetErrorString(e));
        cmgpu_ctx_destroy(c);
        return nullptr;
    };
    cudaError_t e;
    if ((e = cudaSetDevice(device)) != cudaSuccess)
        return bail("cudaSetDevice", e);
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess)
        return bail("cudaGetDeviceProperties", e);
    c->num_sms = prop.multiProcessorCount;
    const size_t ring = c->slot_bytes * ring_slots;
    if ((e = cudaMalloc(&c->d_in, ring)) != cudaSuccess)
        return bail("cudaMalloc(ring)", e);
    const size_t ring_out = out_channels ? c->slot_bytes_out * ring_slots : ring;
    if (((flags & CMGPU_SEPARATE_OUT) || out_channels) && (e = cudaMalloc(&c->d_out, ring_out)) != cudaSuccess)
        return bail("cudaMalloc(out ring)", e);
    if ((e = cudaMemset(c->d_in, 0, ring)) != cudaSuccess ||
        (c->d_out && (e = cudaMemset(c->d_out, 0, ring_out)) != cudaSuccess))
        return bail("cudaMemset", e);
    if (out_channels) {
        if (!(flags & CMGPU_NO_PINNED) && (e = cudaMallocHost(&c->h_ring_out, ring_out)) != cudaSuccess)
            return bail("cudaMallocHost(out staging)", e);
        if (c->h_ring_out)
            memset(c->h_ring_out, 0, ring_out);
        if ((e = cudaMalloc(&c->d_mix, sizeof(MixRow) * max_streams)) != cudaSuccess ||
            (e = cudaMemset(c->d_mix, 0, sizeof(MixRow) * max_streams)) != cudaSuccess)
            return bail("cudaMalloc(mix rows)", e);
        if ((e = cudaMalloc(&c->d_meters_in, sizeof(uint64_t) * c->row_in_u64 * max_streams)) != cudaSuccess ||
            (e = cudaMemset(c->d_meters_in, 0, sizeof(uint64_t) * c->row_in_u64 * max_streams)) != cudaSuccess)
            return bail("cudaMalloc(input meters)", e);
        c->h_mix.assign(max_streams, MixRow());
        for (auto &r : c->h_mix) {
            memset(&r, 0, sizeof(r));
            r.magic = 0x80000001u;      // scale 1: M = 2^31 + 1, shift 31; all-zero weights = silence
            r.shift = 31;
        }
        c->mix_dirty = true;
    }
    if (!(flags & CMGPU_NO_PINNED) && (e = cudaMallocHost(&c->h_ring, ring)) != cudaSuccess)
        return bail("cudaMallocHost(staging)", e);
    if (c->h_ring)
        memset(c->h_ring, 0, ring);
    if ((flags & CMGPU_PLANAR_F32) && !out_channels) {
        c->plane_stride = ((size_t)block_frames + 3u) & ~(size_t)3u;
        c->planar_slot_floats = c->plane_stride * channels * max_streams;
        if ((e = cudaMalloc(&c->d_planar, c->planar_slot_floats * ring_slots * sizeof(float))) != cudaSuccess ||
            (e = cudaMemset(c->d_planar, 0, c->planar_slot_floats * ring_slots * sizeof(float))) != cudaSuccess)
            return bail("cudaMalloc(planes)", e);
    }
    if ((e = cudaMalloc(&c->d_gains, sizeof(GainRow) * max_streams)) != cudaSuccess)
        return bail("cudaMalloc(gains)", e);
    if ((e = cudaMalloc(&c->d_meters, sizeof(uint64_t) * c->row_u64 * max_streams)) != cudaSuccess)
        return bail("cudaMalloc(meters)", e);
    if ((e = cudaMemset(c->d_meters, 0, sizeof(uint64_t) * c->row_u64 * max_streams)) != cudaSuccess)
        return bail("cudaMemset(meters)", e);
    if ((e = cudaMalloc(&c->d_tick, 2 * sizeof(unsigned long long))) != cudaSuccess ||
        (e = cudaMemset(c->d_tick, 0, 2 * sizeof(unsigned long long))) != cudaSuccess)
        return bail("cudaMalloc(tick)", e);
    if ((e = cudaMalloc(&c->d_frames, sizeof(uint32_t) * (size_t)max_streams * ring_slots)) != cudaSuccess)
        return bail("cudaMalloc(frames)", e);
    if ((e = cudaMalloc(&c->d_work, sizeof(unsigned int) * cmgpu_ctx::kWorkCounters)) != cudaSuccess ||
        (e = cudaMemset(c->d_work, 0, sizeof(unsigned int) * cmgpu_ctx::kWorkCounters)) != cudaSuccess)
        return bail("cudaMalloc(work counters)", e);
    if ((e = cudaMalloc(&c->d_done_count, sizeof(unsigned int))) != cudaSuccess ||
        (e = cudaMemset(c->d_done_count, 0, sizeof(unsigned int))) != cudaSuccess)
        return bail("cudaMalloc(completion count)", e);
    {
        // the completion word is an optimisation: without mapped host memory the waits ask the driver
        void *h = nullptr, *d = nullptr;
        if (!getenv("CMGPU_NO_DONE_WORD") && cudaHostAlloc(&h, sizeof(unsigned int), cudaHostAllocMapped) == cudaSuccess) {
            if (cudaHostGetDevicePointer(&d, h, 0) == cudaSuccess) {
                c->h_done = static_cast<volatile unsigned int *>(h);
                *c->h_done = 0;
                c->d_done_flag = static_cast<unsigned int *>(d);
            } else {
                cudaFreeHost(h);
            }
        }
        cudaGetLastError();
    }
    if ((e = cudaStreamCreateWithFlags(&c->s_up, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&c->s_cmp, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&c->s_down, cudaStreamNonBlocking)) != cudaSuccess)
        return bail("cudaStreamCreate", e);
    c->ev_up.assign(ring_slots, nullptr);
    c->ev_cmp.assign(ring_slots, nullptr);
    c->ev_down.assign(ring_slots, nullptr);
    c->up_pending.assign(ring_slots, 0);
    c->down_pending.assign(ring_slots, 0);
    c->cmp_unrecorded.assign(ring_slots, 0);
    c->from_staging.assign(ring_slots, 0);
    c->slot_dirty.assign(ring_slots, 1);
    c->in_chain.assign(ring_slots, 0);
    for (unsigned i = 0; i < ring_slots; i++) {
        if ((e = cudaEventCreateWithFlags(&c->ev_up[i], cudaEventDisableTiming)) != cudaSuccess ||
            (e = cudaEventCreateWithFlags(&c->ev_cmp[i], cudaEventDisableTiming)) != cudaSuccess ||
            (e = cudaEventCreateWithFlags(&c->ev_down[i], cudaEventDisableTiming)) != cudaSuccess)
            return bail("cudaEventCreate", e);
    }
    if ((e = cudaEventCreate(&c->ev_t0)) != cudaSuccess || (e = cudaEventCreate(&c->ev_t1)) != cudaSuccess)
        return bail("cudaEventCreate", e);

    c->has_frames.assign(ring_slots, 0);
    c->h_gains.resize(max_streams);
    c->h_scale.assign(max_streams, 0);
    c->h_gain.assign((size_t)max_streams * channels, 0);
    for (unsigned s = 0; s < max_streams; s++)
        set_row(c, s, 0, nullptr);
    c->dirty_lo = 0;
    c->dirty_hi = max_streams;
    make_plan(c);
    if ((uint64_t)max_streams * c->plan_items >= 0xffffffffull) {
        fail(CMGPU_ERR_INVAL, "cmgpu_ctx_create: %u streams x %u work items per stream-block exceed 2^32", max_streams,
             c->plan_items);
        cmgpu_ctx_destroy(c);
        return nullptr;
    }
    return c;
}

cmgpu_ctx_t *cmgpu_ctx_create(int device, unsigned channels, unsigned max_streams, unsigned ring_slots,
                              unsigned block_frames, unsigned flags)
{
    return ctx_create_impl(device, channels, 0, max_streams, ring_slots, block_frames, flags);
}

cmgpu_ctx_t *cmgpu_mix_ctx_create(int device, unsigned in_channels, unsigned out_channels, unsigned max_streams,
                                  unsigned ring_slots, unsigned block_frames, unsigned flags)
{
    if (!out_channels) {
        fail(CMGPU_ERR_INVAL, "cmgpu_mix_ctx_create: out_channels must be 1..16");
        return nullptr;
    }
    return ctx_create_impl(device, in_channels, out_channels, max_streams, ring_slots, block_frames, flags);
}

void cmgpu_ctx_destroy(cmgpu_ctx_t *c)
{
    if (!c)
        return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    for (auto ev : c->ev_up) if (ev) cudaEventDestroy(ev);
    for (auto ev : c->ev_cmp) if (ev) cudaEventDestroy(ev);
    for (auto ev : c->ev_down) if (ev) cudaEventDestroy(ev);
    if (c->graph) cudaGraphExecDestroy(c->graph);
    for (auto st : c->s_fork) if (st) cudaStreamDestroy(st);
    for (auto ev : c->ev_join) if (ev) cudaEventDestroy(ev);
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_t0) cudaEventDestroy(c->ev_t0);
    if (c->ev_t1) cudaEventDestroy(c->ev_t1);
    cudaFree(c->d_done_count);
    if (c->h_done) cudaFreeHost(const_cast<unsigned int *>(c->h_done));
    if (c->s_up) cudaStreamDestroy(c->s_up);
    if (c->s_cmp) cudaStreamDestroy(c->s_cmp);
    if (c->s_down) cudaStreamDestroy(c->s_down);
    cudaFree(c->d_in);
    cudaFree(c->d_out);
    cudaFree(c->d_gains);
    cudaFree(c->d_meters);
    cudaFree(c->d_frames);
    cudaFree(c->d_work);
    cudaFree(c->d_planar);
    cudaFree(c->d_mix);
    cudaFree(c->d_meters_in);
    if (c->h_ring_out)
        cudaFreeHost(c->h_ring_out);
    cudaFree(c->d_tick);
    if (c->h_ring)
        cudaFreeHost(c->h_ring);
    cudaFree(c->d_take);
    cudaFree(c->d_results);
    cudaFree(c->d_colors);
    cudaFree(c->d_tone);
    if (c->h_take)
        cudaFreeHost(c->h_take);
    if (c->h_frames)
        cudaFreeHost(c->h_frames);
    for (auto ev : c->ev_frames) if (ev) cudaEventDestroy(ev);
    cudaGetLastError();
    delete c;
}

unsigned cmgpu_channels(const cmgpu_ctx_t *c) { return c ? c->channels : 0; }
unsigned cmgpu_max_streams(const cmgpu_ctx_t *c) { return c ? c->max_streams : 0; }
unsigned cmgpu_ring_slots(const cmgpu_ctx_t *c) { return c ? c->slots : 0; }
unsigned cmgpu_block_frames(const cmgpu_ctx_t *c) { return c ? c->block_frames : 0; }
size_t cmgpu_block_stride(const cmgpu_ctx_t *c) { return c ? c->stride : 0; }
size_t cmgpu_slot_bytes(const cmgpu_ctx_t *c) { return c ? c->slot_bytes : 0; }
uint64_t cmgpu_launch_count(const cmgpu_ctx_t *c) { return c ? c->launches : 0; }
uint64_t cmgpu_word_waits(const cmgpu_ctx_t *c) { return c ? c->word_waits.load(std::memory_order_relaxed) : 0; }
int cmgpu_transfer_bytes(const cmgpu_ctx_t *c, uint64_t *h2d, uint64_t *d2h)
{
    if (!c)
        return fail(CMGPU_ERR_FAULT, "NULL context");
    if (h2d)
        *h2d = c->bytes_h2d;
    if (d2h)
        *d2h = c->bytes_d2h;
    return CMGPU_OK;
}
const char *cmgpu_kernel_name(const cmgpu_ctx_t *c)
{
    if (!c)
        return "";
    if (c->out_channels)
        return (c->channels == 8 && c->out_channels == 2 && !(c->flags & CMGPU_FORCE_GENERIC))
                   ? ((c->flags & CMGPU_MIX_OUTPUT_METER_ONLY) ? "mix8to2_tick<outputs metered>" : "mix8to2_tick")
                   : "mix_tick<generic>";
    return c->kname_last[0] ? c->kname_last : c->kname;       // (the span kernel is chosen per launch, not per context)
}
unsigned cmgpu_meter_row_u64(const cmgpu_ctx_t *c) { return c ? c->row_u64 : 0; }
void *cmgpu_device_meters(cmgpu_ctx_t *c) { return c ? c->d_meters : nullptr; }

int cmgpu_set_active_streams(cmgpu_ctx_t *c, unsigned n)
{
    if (!c)
        return fail(CMGPU_ERR_FAULT, "NULL context");
    if (n > c->max_streams)
        return fail(CMGPU_ERR_INVAL, "active streams %u > max_streams %u", n, c->max_streams);
    std::lock_guard<std::mutex> lk(c->mu);
    c->active = n;
    c->classes_dirty = true;
    c->config_gen++;
    return CMGPU_OK;
}

int cmgpu_stream_set_gain(cmgpu_ctx_t *c, unsigned stream, unsigned n, uint16_t scale, const uint16_t *gain)
{
    if (!c)
        return fail(CMGPU_ERR_FAULT, "NULL context");
    if (stream >= c->max_streams)
        return fail(CMGPU_ERR_INVAL, "stream %u out of range", stream);
    std::lock_guard<std::mutex> lk(c->mu);
    uint16_t g[CMGPU_MAX_CHANNELS];
    // transform.c:200-221
    if (!n || !scale || !gain) {
        set_row(c, stream, 0, nullptr);
    } else if (n == c->channels) {
        memcpy(g, gain, sizeof(uint16_t) * n);
        set_row(c, stream, scale, g);
    } else if (n == 1) {
        for (unsigned ch = 0; ch < c->channels; ch++)
            g[ch] = gain[0];
        set_row(c, stream, scale, g);
    } else if (n == 2 && c->channels == 1) {
        g[0] = (uint16_t)(((uint32_t)gain[0] + (uint32_t)gain[1]) / 2u);
        set_row(c, stream, scale, g);
    } else {
        return fail(CMGPU_ERR_INVAL, "gain for %u channels cannot be mapped onto %u", n, c->channels);
    }
    mark_dirty(c, stream, stream + 1);
    return CMGPU_OK;
}

int cmgpu_set_gain_table(cmgpu_ctx_t *c, unsigned first, unsigned count, const uint16_t *scale, const uint16_t *gain)
{
    if (!c || !scale || !gain)
        return fail(CMGPU_ERR_FAULT, "NULL argument");
    if ((uint64_t)first + count > c->max_streams)
        return fail(CMGPU_ERR_INVAL, "stream range out of bounds");
    std::lock_guard<std::mutex> lk(c->mu);
    for (unsigned i = 0; i < count; i++)
        set_row(c, first + i, scale[i], gain + (size_t)i * c->channels);
    if (count)
        mark_dirty(c, first, first + count);
    return CMGPU_OK;
}

int cmgpu_stream_get_gain(const cmgpu_ctx_t *c, unsigned stream, uint16_t *scale, uint16_t gain[CMGPU_MAX_CHANNELS])
{
    if (!c || !scale || !gain)
        return fail(CMGPU_ERR_FAULT, "NULL argument");
    if (stream >= c->max_streams)
        return fail(CMGPU_ERR_INVAL, "stream %u out of range", stream);
    *scale = c->h_scale[stream];
    for (unsigned ch = 0; ch < CMGPU_MAX_CHANNELS; ch++)
        gain[ch] = ch < c->channels ? c->h_gain[(size_t)stream * c->channels + ch] : 0;
    return CMGPU_OK;
}

void *cmgpu_host_slot(cmgpu_ctx_t *c, unsigned slot)
{
    return slot_ok(c, slot) && c->h_ring ? c->h_ring + (size_t)slot * c->slot_bytes : nullptr;
}
void *cmgpu_device_slot(cmgpu_ctx_t *c, unsigned slot)
{
    return slot_ok(c, slot) ? c->d_in + (size_t)slot * c->slot_bytes : nullptr;
}
void *cmgpu_device_out_slot(cmgpu_ctx_t *c, unsigned slot)
{
    if (!slot_ok(c, slot))
        return nullptr;
    return (c->d_out ? c->d_out : c->d_in) + (size_t)slot * (c->out_channels ? c->slot_bytes_out : c->slot_bytes);
}
void *cmgpu_host_out_slot(cmgpu_ctx_t *c, unsigned slot)
{
    if (!slot_ok(c, slot))
        return nullptr;
    if (c->out_channels)
        return c->h_ring_out ? c->h_ring_out + (size_t)slot * c->slot_bytes_out : nullptr;
    return c->h_ring ? c->h_ring + (size_t)slot * c->slot_bytes : nullptr;
}
unsigned cmgpu_out_channels(const cmgpu_ctx_t *c) { return c ? (c->out_channels ? c->out_channels : c->channels) : 0; }
size_t cmgpu_out_block_stride(const cmgpu_ctx_t *c) { return c ? (c->out_channels ? c->stride_out : c->stride) : 0; }

int cmgpu_slot_set_frames(cmgpu_ctx_t *c, unsigned slot, const uint32_t *frames)
{
    if (!slot_ok(c, slot))
        return fail(c ? CMGPU_ERR_INVAL : CMGPU_ERR_FAULT, "bad context or slot");
    std::lock_guard<std::mutex> lk(c->mu);
    c->config_gen++;
    if (!frames) {
        c->has_frames[slot] = 0;
        return CMGPU_OK;
    }
    for (unsigned s = 0; s < c->active; s++)
        if (frames[s] > c->block_frames)
            return fail(CMGPU_ERR_INVAL, "stream %u: %u frames > block_frames %u", s, frames[s], c->block_frames);
    CU(cudaSetDevice(c->device));
    // "Copied": the counts are staged in a context-owned pinned row of the slot, so the caller's array
    // is free on return whether or not it is page-locked; the row is reused only after its previous
    // upload has left it
    if (!c->h_frames) {
        CU(cudaMallocHost(&c->h_frames, sizeof(uint32_t) * (size_t)c->max_streams * c->slots));
        c->ev_frames.assign(c->slots, nullptr);
        for (unsigned i = 0; i < c->slots; i++)
            CU(cudaEventCreateWithFlags(&c->ev_frames[i], cudaEventDisableTiming));
    }
    CU(cudaEventSynchronize(c->ev_frames[slot]));
    uint32_t *stage = c->h_frames + (size_t)slot * c->max_streams;
    memcpy(stage, frames, sizeof(uint32_t) * c->active);
    CU(cudaMemcpyAsync(c->d_frames + (size_t)slot * c->max_streams, stage, sizeof(uint32_t) * c->active,
                       cudaMemcpyHostToDevice, c->cmp()));
    CU(cudaEventRecord(c->ev_frames[slot], c->cmp()));
    c->has_frames[slot] = 1;
    return CMGPU_OK;
}

int cmgpu_submit(cmgpu_ctx_t *c, unsigned slot, const void *host)
{
    CMGPU_TRACE("cmgpu_submit");
    if (!slot_ok(c, slot))
        return fail(c ? CMGPU_ERR_INVAL : CMGPU_ERR_FAULT, "bad context or slot");
    std::lock_guard<std::mutex> lk(c->mu);
    if (!host)
        host = c->h_ring ? c->h_ring + (size_t)slot * c->slot_bytes : nullptr;
    if (!host)
        return fail(CMGPU_ERR_FAULT, "no host buffer and no pinned staging");
    CU(cudaSetDevice(c->device));
    // the slot must not be overwritten while its previous tick or download is in flight
    if (int erc = ticks_done_event_locked(c, slot))
        return erc;
    c->up_seq.fetch_add(1, std::memory_order_release);
    CU(cudaStreamWaitEvent(c->s_up, c->ev_cmp[slot], 0));
    CU(cudaStreamWaitEvent(c->s_up, c->ev_down[slot], 0));
    CU(cudaMemcpyAsync(c->d_in + (size_t)slot * c->slot_bytes, host, c->stride * c->active, cudaMemcpyHostToDevice,
                       c->s_up));
    CU(cudaEventRecord(c->ev_up[slot], c->s_up));
    c->up_pending[slot] = 1;
    c->from_staging[slot] = c->h_ring && host == c->h_ring + (size_t)slot * c->slot_bytes;
    c->slot_dirty[slot] = 0;
    c->bytes_h2d += c->stride * c->active;
    return CMGPU_OK;
}

int cmgpu_process(cmgpu_ctx_t *c, unsigned slot, unsigned flags)
{
    CMGPU_TRACE("cmgpu_process");
    if (!slot_ok(c, slot))
        return fail(c ? CMGPU_ERR_INVAL : CMGPU_ERR_FAULT, "bad context or slot");
    std::lock_guard<std::mutex> lk(c->mu);
    CU(cudaSetDevice(c->device));
    // a tick that consumes a fresh upload (or follows a download) is part of a streaming pipeline: the
    // slot's next upload will want to know exactly when THIS tick is done, so say so now; a tick on
    // resident data records nothing and stays a bare kernel launch
    const bool streaming = c->up_pending[slot] || c->down_pending[slot];
    int rc = tick_waits_locked(c, slot);
    if (rc || (rc = launch_locked(c, slot, flags)))
        return rc;
    c->cmp_unrecorded[slot] = 1;
    if (streaming)
        rc = ticks_done_event_locked(c, slot);
    return rc;
}

int cmgpu_fetch(cmgpu_ctx_t *c, unsigned slot, void *host)
{
    CMGPU_TRACE("cmgpu_fetch");
    if (!slot_ok(c, slot))
        return fail(c ? CMGPU_ERR_INVAL : CMGPU_ERR_FAULT, "bad context or slot");
    std::lock_guard<std::mutex> lk(c->mu);
    const size_t out_slot = c->out_channels ? c->slot_bytes_out : c->slot_bytes;
    const size_t out_stride = c->out_channels ? c->stride_out : c->stride;
    uint8_t *staging = c->out_channels ? c->h_ring_out : c->h_ring;
    if (!host)
        host = staging ? staging + (size_t)slot * out_slot : nullptr;
    if (!host)
        return fail(CMGPU_ERR_FAULT, "no host buffer and no pinned staging");
    CU(cudaSetDevice(c->device));
    // nothing on the device differs from what the staging slot already holds (in place, no tick has
    // written PCM since the upload from that very slot): the download would copy the input onto itself
    if (!c->d_out && !c->out_channels && staging && host == staging + (size_t)slot * out_slot && c->from_staging[slot] &&
        !c->slot_dirty[slot])
        return CMGPU_OK;
    if (int erc = ticks_done_event_locked(c, slot))
        return erc;
    c->down_seq.fetch_add(1, std::memory_order_release);
    CU(cudaStreamWaitEvent(c->s_down, c->ev_cmp[slot], 0));
    CU(cudaStreamWaitEvent(c->s_down, c->ev_up[slot], 0));
    c->bytes_d2h += out_stride * c->active;
    const uint8_t *src = (c->d_out ? c->d_out : c->d_in) + (size_t)slot * out_slot;
    CU(cudaMemcpyAsync(host, src, out_stride * c->active, cudaMemcpyDeviceToHost, c->s_down));
    CU(cudaEventRecord(c->ev_down[slot], c->s_down));
    c->down_pending[slot] = 1;
    return CMGPU_OK;
}

void *cmgpu_device_planar_slot(cmgpu_ctx_t *c, unsigned slot)
{
    return slot_ok(c, slot) && c->d_planar ? c->d_planar + (size_t)slot * c->planar_slot_floats : nullptr;
}
size_t cmgpu_plane_stride(const cmgpu_ctx_t *c) { return c ? c->plane_stride : 0; }

int cmgpu_fetch_planar(cmgpu_ctx_t *c, unsigned slot, float *host)
{
    if (!slot_ok(c, slot))
        return fail(c ? CMGPU_ERR_INVAL : CMGPU_ERR_FAULT, "bad context or slot");
    if (!host)
        return fail(CMGPU_ERR_FAULT, "NULL host buffer");
    if (!c->d_planar)
        return fail(CMGPU_ERR_INVAL, "context has no float planes (CMGPU_PLANAR_F32)");
    std::lock_guard<std::mutex> lk(c->mu);
    CU(cudaSetDevice(c->device));
    if (int erc = ticks_done_event_locked(c, slot))      // a tick on resident data has not recorded it yet
        return erc;
    c->down_seq.fetch_add(1, std::memory_order_release);
    CU(cudaStreamWaitEvent(c->s_down, c->ev_cmp[slot], 0));
    CU(cudaMemcpyAsync(host, c->d_planar + (size_t)slot * c->planar_slot_floats,
                       c->plane_stride * c->channels * c->active * sizeof(float), cudaMemcpyDeviceToHost, c->s_down));
    CU(cudaEventRecord(c->ev_down[slot], c->s_down));
    c->down_pending[slot] = 1;
    return CMGPU_OK;
}

int cmgpu_debug_violations(void)
{
#ifdef CMGPU_BOUNDS_CHECK
    unsigned int n = 0;
    if (cudaMemcpyFromSymbol(&n, cmgpu::g_dbg_violations, sizeof(n)) != cudaSuccess)
        return -1;
    return (int)(n > 0x7fffffffu ? 0x7fffffffu : n);
#else
    return -1;          // not a bounds-checking build
#endif
}

// Waiting for a stream or an event: poll for a short while before blocking. A blocking wait costs a
// few microseconds of wake-up latency, which is most of what is left of a 20 ms-sized tick issued
// alone (config 3: ~8 us on the device); work that is not about to finish falls through to the
// blocking call after kSpinNs and costs nothing extra.
namespace {
constexpr long long kSpinNs = 60000;
inline long long now_ns()
{
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (long long)ts.tv_sec * 1000000000ll + ts.tv_nsec;
}
cudaError_t wait_stream(cudaStream_t st)
{
    const long long t0 = now_ns();
    for (;;) {
        const cudaError_t e = cudaStreamQuery(st);
        if (e != cudaErrorNotReady)
            return e;
        if (now_ns() - t0 > kSpinNs)
            return cudaStreamSynchronize(st);
    }
}
cudaError_t wait_event(cudaEvent_t ev)
{
    const long long t0 = now_ns();
    for (;;) {
        const cudaError_t e = cudaEventQuery(ev);
        if (e != cudaErrorNotReady)
            return e;
        if (now_ns() - t0 > kSpinNs)
            return cudaEventSynchronize(ev);
    }
}
}  // namespace

int cmgpu_sync(cmgpu_ctx_t *c)
{
    CMGPU_TRACE("cmgpu_sync");
    if (!c)
        return fail(CMGPU_ERR_FAULT, "NULL context");
    CU(cudaSetDevice(c->device));
    // side streams: only if something was queued on them since the last wait (a query of an idle stream
    // still costs 1.4 us)
    const uint64_t up = c->up_seq.load(std::memory_order_acquire), down = c->down_seq.load(std::memory_order_acquire);
    if (up != c->up_synced.load(std::memory_order_relaxed)) {
        CU(wait_stream(c->s_up));
        c->up_synced.store(up, std::memory_order_relaxed);
    }
    // compute stream: when its tail is a tick launch, that launch's completion word says when it is done
    // (and, launches completing in stream order, everything before it); a launch that never finishes --
    // a fault -- never writes the word, and the blocking wait below reports the error
    bool cmp_done = false;
    if (const uint32_t gen = c->tail_gen.load(std::memory_order_acquire)) {
        const long long t0 = now_ns();
        unsigned spins = 0;
        // (>= in generation order: another thread's later launch may have finished as well by now)
        while (!(cmp_done = (int32_t)(*c->h_done - gen) >= 0))
            if ((++spins & 63u) == 0 && now_ns() - t0 > kSpinNs)
                break;
    }
    if (cmp_done)
        c->word_waits.fetch_add(1, std::memory_order_relaxed);
    else
        CU(wait_stream(c->s_cmp));
    if (down != c->down_synced.load(std::memory_order_relaxed)) {
        CU(wait_stream(c->s_down));
        c->down_synced.store(down, std::memory_order_relaxed);
    }
    c->idle_hint.store(true, std::memory_order_release);       // the next tick launch is waited for tick by tick, it seems
#ifdef CMGPU_BOUNDS_CHECK
    if (const int n = cmgpu_debug_violations())
        return fail(CMGPU_ERR_GENERIC, "bounds check: %d PCM vector accesses outside the context's rings", n);
#endif
    return CMGPU_OK;
}

int cmgpu_slot_wait(cmgpu_ctx_t *c, unsigned slot)
{
    if (!slot_ok(c, slot))
        return fail(c ? CMGPU_ERR_INVAL : CMGPU_ERR_FAULT, "bad context or slot");
    CU(cudaSetDevice(c->device));
    {
        std::lock_guard<std::mutex> lk(c->mu);           // chain / event bookkeeping only; the waits run unlocked
        if (int erc = ticks_done_event_locked(c, slot))
            return erc;
    }
    CU(wait_event(c->ev_up[slot]));
    CU(wait_event(c->ev_cmp[slot]));
    CU(wait_event(c->ev_down[slot]));
    return CMGPU_OK;
}

int cmgpu_meter_decode(const uint64_t *rows, unsigned count, unsigned channels, cmgpu_meter_state_t *out)
{
    if (!rows || !out)
        return fail(CMGPU_ERR_FAULT, "NULL argument");
    if (!channels || channels > CMGPU_MAX_CHANNELS)
        return fail(CMGPU_ERR_INVAL, "channels out of range");
    const unsigned row = 2 * channels + 2;
    for (unsigned i = 0; i < count; i++)
        cmgpu::decode_row(rows + (size_t)i * row, channels, out + i);
    return CMGPU_OK;
}

int cmgpu_meter_snapshot(cmgpu_ctx_t *c, unsigned first, unsigned count, cmgpu_meter_state_t *out, int reset)
{
    CMGPU_TRACE("cmgpu_meter_snapshot");
    if (!c || !out)
        return fail(CMGPU_ERR_FAULT, "NULL argument");
    if ((uint64_t)first + count > c->max_streams)
        return fail(CMGPU_ERR_INVAL, "stream range out of bounds");
    if (!count)
        return CMGPU_OK;
    std::lock_guard<std::mutex> lk(c->mu);
    CU(cudaSetDevice(c->device));
    const size_t n = (size_t)count * c->row_u64;
    c->scratch.resize(n);
    unsigned long long *src = c->d_meters + (size_t)first * c->row_u64;
    CU(cudaMemcpyAsync(c->scratch.data(), src, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, c->cmp()));
    if (reset) {
        CU(cudaMemsetAsync(src, 0, n * sizeof(uint64_t), c->cmp()));
        if (first == 0 && count == c->max_streams && !c->out_channels) {      // as in cmgpu_meter_reset
            CU(cudaMemsetAsync(c->d_tick, 0, sizeof(unsigned long long), c->cmp()));
            c->pending_ticks = 0;
        }
    }
    c->last_first = ~0u;
    c->chain_open = false;
    CU(cudaStreamSynchronize(c->cmp()));
    for (unsigned i = 0; i < count; i++)
        cmgpu::decode_row(c->scratch.data() + (size_t)i * c->row_u64, c->out_channels ? c->out_channels : c->channels, out + i);
    return CMGPU_OK;
}

int cmgpu_mix_input_snapshot(cmgpu_ctx_t *c, unsigned first, unsigned count, cmgpu_meter_state_t *out, int reset)
{
    if (!c || !out)
        return fail(CMGPU_ERR_FAULT, "NULL argument");
    if (!c->out_channels)
        return fail(CMGPU_ERR_INVAL, "not a downmix context");
    if ((uint64_t)first + count > c->max_streams)
        return fail(CMGPU_ERR_INVAL, "stream range out of bounds");
    if (!count)
        return CMGPU_OK;
    std::lock_guard<std::mutex> lk(c->mu);
    CU(cudaSetDevice(c->device));
    const size_t n = (size_t)count * c->row_in_u64;
    c->scratch.resize(n);
    unsigned long long *src = c->d_meters_in + (size_t)first * c->row_in_u64;
    CU(cudaMemcpyAsync(c->scratch.data(), src, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, c->cmp()));
    if (reset)
        CU(cudaMemsetAsync(src, 0, n * sizeof(uint64_t), c->cmp()));
    CU(cudaStreamSynchronize(c->cmp()));
    for (unsigned i = 0; i < count; i++)
        cmgpu::decode_row(c->scratch.data() + (size_t)i * c->row_in_u64, c->channels, out + i);
    return CMGPU_OK;
}

int cmgpu_stream_set_mix(cmgpu_ctx_t *c, unsigned stream, uint16_t scale, const uint16_t *weights)
{
    if (!c || !weights)
        return fail(CMGPU_ERR_FAULT, "NULL argument");
    if (!c->out_channels)
        return fail(CMGPU_ERR_INVAL, "not a downmix context");
    if (stream >= c->max_streams || !scale)
        return fail(CMGPU_ERR_INVAL, "stream out of range or scale 0");
    std::lock_guard<std::mutex> lk(c->mu);
    MixRow &r = c->h_mix[stream];
    memset(&r, 0, sizeof(r));
    for (unsigned m = 0; m < c->out_channels; m++)
        for (unsigned ch = 0; ch < c->channels; ch++)
            r.w[m][ch] = weights[(size_t)m * c->channels + ch];
    if (c->channels == 8 && c->out_channels == 2) {
        for (unsigned m = 0; m < 2; m++)
            for (unsigned p = 0; p < 4; p++) {
                const uint32_t w0 = r.w[m][2 * p], w1 = r.w[m][2 * p + 1];
                r.packed[m][p] = (w0 & 0xffu) | ((w1 & 0xffu) << 8) | ((w0 >> 8) << 16) | ((w1 >> 8) << 24);
            }
    }
    const unsigned l = ceil_log2(scale);
    r.shift = 31 + l;
    r.magic = (uint32_t)((((uint64_t)1 << r.shift) / scale) + 1);     // < 2^32: 2^(31+l)/scale < 2^32 for scale > 2^(l-1)
    c->mix_dirty = true;
    c->config_gen++;
    return CMGPU_OK;
}

int cmgpu_meter_reset(cmgpu_ctx_t *c, unsigned first, unsigned count)
{
    if (!c)
        return fail(CMGPU_ERR_FAULT, "NULL context");
    if ((uint64_t)first + count > c->max_streams)
        return fail(CMGPU_ERR_INVAL, "stream range out of bounds");
    if (!count)
        return CMGPU_OK;
    std::lock_guard<std::mutex> lk(c->mu);
    CU(cudaSetDevice(c->device));
    CU(cudaMemsetAsync(c->d_meters + (size_t)first * c->row_u64, 0, sizeof(uint64_t) * c->row_u64 * count, c->cmp()));
    if (c->d_meters_in)
        CU(cudaMemsetAsync(c->d_meters_in + (size_t)first * c->row_in_u64, 0, sizeof(uint64_t) * c->row_in_u64 * count,
                           c->cmp()));
    if (first == 0 && count == c->max_streams) {
        // every meter window starts afresh: rebase the position keys' tick number to zero, so that the
        // 46 - pbits bits a key keeps of it can only wrap INSIDE one window of that many ticks
        CU(cudaMemsetAsync(c->d_tick, 0, sizeof(unsigned long long), c->cmp()));
        c->pending_ticks = 0;
    }
    c->last_first = ~0u;
    c->chain_open = false;
    return CMGPU_OK;
}

int cmgpu_finalise(const cmgpu_meter_state_t *st, uint32_t rate, unsigned channels, cmgpu_result_t *out)
{
    if (!st || !out)
        return fail(CMGPU_ERR_FAULT, "NULL argument");
    if (!channels || channels > CMGPU_MAX_CHANNELS)
        return fail(CMGPU_ERR_INVAL, "channels out of range");
    return cmgpu::finalise_state(st, rate, channels, out);
}

int cmgpu_time_process(cmgpu_ctx_t *c, unsigned first_slot, unsigned n_slots, unsigned reps, unsigned flags, float *ms)
{
    if (!c || !ms)
        return fail(CMGPU_ERR_FAULT, "NULL argument");
    if (!n_slots || (uint64_t)first_slot + n_slots > c->slots)
        return fail(CMGPU_ERR_INVAL, "slot range out of bounds");
    std::lock_guard<std::mutex> lk(c->mu);
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->s_up));
    CU(cudaStreamSynchronize(c->s_down));
    int rc = upload_gains_locked(c);
    if (rc)
        return rc;
    CU(cudaEventRecord(c->ev_t0, c->cmp()));
    for (unsigned r = 0; r < reps; r++) {
        rc = launch_locked(c, first_slot + r % n_slots, flags);
        if (rc)
            return rc;
    }
    CU(cudaEventRecord(c->ev_t1, c->cmp()));
    CU(cudaEventSynchronize(c->ev_t1));
    CU(cudaEventElapsedTime(ms, c->ev_t0, c->ev_t1));
    return CMGPU_OK;
}

// A cycle of the vector kernels needs no graph at all: the ring is one allocation, so ONE launch can
// walk all its slots as a span (work items numbered (tick, stream, chunk), positions ordered by
// tick). That turns the launch-bound small-buffer regime into the same software-pipelined stream of
// items as a large tick. Needs the fast kernels, no float planes, and per-slot frame counts given
// for all of the span's slots or for none.
// Brings the device's tick counter up to date with the ticks the host has numbered itself; needed
// before a captured cycle, whose launches count from the device counter.
}  // extern "C"

namespace {
int flush_ticks_locked(cmgpu_ctx *c)
{
    c->last_first = ~0u;
    c->chain_open = false;
    if (!c->pending_ticks)
        return CMGPU_OK;
    cmgpu::bump_tick<<<1, 32, 0, c->cmp()>>>(c->d_tick, c->pending_ticks);
    CU(cudaGetLastError());
    c->pending_ticks = 0;
    return CMGPU_OK;
}

static bool span_ok(const cmgpu_ctx *c, unsigned first_slot, unsigned n_slots, unsigned flags)
{
    if (n_slots < 2 || c->plan_g <= 0 || c->tma || c->out_channels || (flags & CMGPU_PLANAR) || c->env_no_span)
        return false;
    if ((uint64_t)n_slots * c->active * c->plan_items >= (1ull << 32))
        return false;
    unsigned with_frames = 0;
    for (unsigned i = 0; i < n_slots; i++)
        with_frames += c->has_frames[first_slot + i] ? 1u : 0u;
    return with_frames == 0 || with_frames == n_slots;
}

static int build_cycle_locked(cmgpu_ctx *c, unsigned first_slot, unsigned n_slots, unsigned flags)
{
    if (c->graph && c->graph_gen == c->config_gen && c->graph_first == first_slot && c->graph_n == n_slots &&
        c->graph_flags == flags)
        return CMGPU_OK;
    if (c->graph) {
        cudaGraphExecDestroy(c->graph);
        c->graph = nullptr;
    }
    // everything a launch may have to upload happens before the capture starts
    int rc = upload_gains_locked(c);
    if (rc)
        return rc;
    CU(cudaStreamSynchronize(c->cmp()));
    const unsigned kFork = 8;
    if (c->s_fork.empty()) {
        c->s_fork.assign(kFork, nullptr);
        c->ev_join.assign(kFork, nullptr);
        for (unsigned i = 0; i < kFork; i++) {
            CU(cudaStreamCreateWithFlags(&c->s_fork[i], cudaStreamNonBlocking));
            CU(cudaEventCreateWithFlags(&c->ev_join[i], cudaEventDisableTiming));
        }
        CU(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    }
    const uint64_t before = c->launches;
    CU(cudaStreamBeginCapture(c->cmp(), cudaStreamCaptureModeThreadLocal));
    const unsigned lanes = n_slots < kFork ? n_slots : kFork;
    cudaError_t ce = cudaEventRecord(c->ev_fork, c->cmp());
    for (unsigned i = 0; i < lanes && ce == cudaSuccess; i++)
        ce = cudaStreamWaitEvent(c->s_fork[i], c->ev_fork, 0);
    for (unsigned i = 0; i < n_slots && rc == CMGPU_OK && ce == cudaSuccess; i++)
        rc = launch_locked(c, first_slot + i, flags, c->s_fork[i % lanes], i, true);
    for (unsigned i = 0; i < lanes && ce == cudaSuccess; i++) {
        ce = cudaEventRecord(c->ev_join[i], c->s_fork[i]);
        if (ce == cudaSuccess)
            ce = cudaStreamWaitEvent(c->cmp(), c->ev_join[i], 0);
    }
    if (ce == cudaSuccess) {
        cmgpu::bump_tick<<<1, 32, 0, c->cmp()>>>(c->d_tick, n_slots);
        ce = cudaGetLastError();             // (the one-thread bump is not counted as a tick launch)
    }
    cudaGraph_t g = nullptr;
    cudaError_t e = cudaStreamEndCapture(c->cmp(), &g);
    c->graph_launches = c->launches - before;
    c->launches = before;                           // capturing is not launching
    if (rc != CMGPU_OK || e != cudaSuccess || ce != cudaSuccess) {
        if (g)
            cudaGraphDestroy(g);
        return rc != CMGPU_OK ? rc
                              : fail(CMGPU_ERR_GENERIC, "capturing the cycle failed: %s",
                                     cudaGetErrorString(e != cudaSuccess ? e : ce));
    }
    e = cudaGraphInstantiate(&c->graph, g, 0);
    cudaGraphDestroy(g);
    if (e != cudaSuccess)
        return fail(CMGPU_ERR_GENERIC, "cudaGraphInstantiate: %s", cudaGetErrorString(e));
    c->graph_first = first_slot;
    c->graph_n = n_slots;
    c->graph_flags = flags;
    c->graph_gen = c->config_gen;
    return CMGPU_OK;
}

}  // namespace

extern "C" {

int cmgpu_process_cycle(cmgpu_ctx_t *c, unsigned first_slot, unsigned n_slots, unsigned flags)
{
    CMGPU_TRACE("cmgpu_process_cycle");
    if (!c)
        return fail(CMGPU_ERR_FAULT, "NULL context");
    if (!n_slots || (uint64_t)first_slot + n_slots > c->slots)
        return fail(CMGPU_ERR_INVAL, "slot range out of bounds");
    std::lock_guard<std::mutex> lk(c->mu);
    CU(cudaSetDevice(c->device));
    const bool span = span_ok(c, first_slot, n_slots, flags);
    int rc = span ? CMGPU_OK : build_cycle_locked(c, first_slot, n_slots, flags);
    if (rc)
        return rc;
    // order the cycle after the uploads of its slots and before their next download
    bool streaming = false;
    for (unsigned i = 0; i < n_slots; i++) {
        streaming = streaming || c->up_pending[first_slot + i] || c->down_pending[first_slot + i];
        if ((rc = tick_waits_locked(c, first_slot + i)))
            return rc;
    }
    if (span) {
        if ((rc = launch_locked(c, first_slot, flags, c->cmp(), 0, false, n_slots)))
            return rc;
    } else {
        if ((rc = flush_ticks_locked(c)))
            return rc;
#ifdef CMGPU_BOUNDS_CHECK
        if ((rc = debug_set_bounds(c, c->cmp())))
            return rc;
#endif
        CU(cudaGraphLaunch(c->graph, c->cmp()));
        c->launches += c->graph_launches;
        for (unsigned i = 0; i < n_slots; i++)
            c->slot_dirty[first_slot + i] = 1;      // (a replayed graph does not say whether it stores: assume it does)
    }
    for (unsigned i = 0; i < n_slots; i++) {
        c->cmp_unrecorded[first_slot + i] = 1;
        if (streaming && (rc = ticks_done_event_locked(c, first_slot + i)))
            return rc;
    }
    return CMGPU_OK;
}

int cmgpu_time_cycles(cmgpu_ctx_t *c, unsigned first_slot, unsigned n_slots, unsigned cycles, unsigned flags, float *ms)
{
    if (!c || !ms)
        return fail(CMGPU_ERR_FAULT, "NULL argument");
    if (!n_slots || (uint64_t)first_slot + n_slots > c->slots)
        return fail(CMGPU_ERR_INVAL, "slot range out of bounds");
    std::lock_guard<std::mutex> lk(c->mu);
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->s_up));
    CU(cudaStreamSynchronize(c->s_down));
    const bool span = span_ok(c, first_slot, n_slots, flags);
    int rc = span ? upload_gains_locked(c) : build_cycle_locked(c, first_slot, n_slots, flags);
    if (rc)
        return rc;
    CU(cudaEventRecord(c->ev_t0, c->cmp()));
    for (unsigned r = 0; r < cycles; r++) {
        if (span) {
            if ((rc = launch_locked(c, first_slot, flags, c->cmp(), 0, false, n_slots)))
                return rc;
        } else {
            if ((rc = flush_ticks_locked(c)))
                return rc;
#ifdef CMGPU_BOUNDS_CHECK
            if ((rc = debug_set_bounds(c, c->cmp())))
                return rc;
#endif
            CU(cudaGraphLaunch(c->graph, c->cmp()));
            c->launches += c->graph_launches;
            for (unsigned i = 0; i < n_slots; i++)
                c->slot_dirty[first_slot + i] = 1;
        }
    }
    CU(cudaEventRecord(c->ev_t1, c->cmp()));
    CU(cudaEventSynchronize(c->ev_t1));
    CU(cudaEventElapsedTime(ms, c->ev_t0, c->ev_t1));
    return CMGPU_OK;
}

int cmgpu_time_single_tick(cmgpu_ctx_t *c, unsigned slot, unsigned flags, unsigned reps, float *median_us, float *min_us)
{
    if (!c || !median_us || !min_us)
        return fail(CMGPU_ERR_FAULT, "NULL argument");
    if (!slot_ok(c, slot) || !reps)
        return fail(CMGPU_ERR_INVAL, "bad slot or reps");
    std::vector<float> us(reps);
    for (unsigned r = 0; r < reps; r++) {
        int rc = cmgpu_sync(c);
        if (rc)
            return rc;
        const long long t0 = now_ns();
        if ((rc = cmgpu_process(c, slot, flags)) || (rc = cmgpu_sync(c)))
            return rc;
        us[r] = (float)(now_ns() - t0) * 1e-3f;
    }
    std::sort(us.begin(), us.end());
    *median_us = us[reps / 2];
    *min_us = us[0];
    return CMGPU_OK;
}

int cmgpu_recipe_eval(uint16_t gain, uint16_t scale, int16_t x)
{
    if (!scale)
        return x;
    return recipe_eval(make_recipe(gain, scale), x);
}

int cmgpu_recipe_table(uint16_t gain, uint16_t scale, int16_t out[65536])
{
    if (!out)
        return fail(CMGPU_ERR_FAULT, "NULL argument");
    RecipeHost r = make_recipe(gain, scale ? scale : 1);
    for (int i = 0; i < 65536; i++)
        out[i] = (int16_t)(scale ? recipe_eval(r, i - 32768) : i - 32768);
    return CMGPU_OK;
}

}  // extern "C"
