// host_common.h -- argument validation, workspace layout and Params construction shared by the
// C-ABI library (tfcfft_api.cu) and the CPU emulation library (emu.cu).
#pragma once
#include <cstdlib>
#include <cstddef>
#include <cstdint>

#include "spectral_core.cuh"
#include "triplet.cuh"
#include "regional.cuh"

namespace tfcfft {

// workspace header (zeroed by tfcfft_workspace_init, left zeroed by every call):
//   [0]    finalise ticket          [64..95] scratch outputs of the spectra entry points     [128] rescale ticket
//   [256]  scheduler of the pipelined sub-tile kernel: 3 queue heads + exit ticket (kSchedHeads)
//   [512]  forward-done counters, one per chunk tile (kSchedMaxTiles)     [512 + 4096] combine-done counters
//   [512 + 8192]  sub-tile path: one byte per forward load unit of the chunk (D*D/2 per tile), 1 = every fake pixel of
//                 the unit equals its real pixel after the luma fold; REWRITTEN by every forward launch before the
//                 combine launch of the same chunk reads it (never relied upon to be zero)
constexpr size_t kWsHeader = 512 + 3 * 4096;
constexpr size_t kSchedHeads = 256, kSchedFwdDone = 512, kSchedCmbDone = 512 + 4096, kEqFlags = 512 + 2 * 4096;
constexpr int kEqFlagBytes = 4096;
constexpr int kSchedMaxTiles = 1024;
constexpr size_t kWsChunkBytes = (size_t)64 << 20;  // spectrum workspace per chunk: stays L2-resident (126 MB L2, pixel streams are evict-first)

inline size_t elem_size(int dtype) {
    switch (dtype) {
        case TFCFFT_F32: return 4;
        case TFCFFT_F16: case TFCFFT_BF16: return 2;
        case TFCFFT_U8: return 1;
        default: return 0;
    }
}
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct Geometry {
    int p;            // patch side
    int cprime;       // spectra per tile position
    bool luma3;       // 3 channels folded to luma
    bool split;       // split path (P = 512, spectra at P >= 256, or forced)
    bool sub;         // sub-tile path (P = 128 / 256): packed 64 x 64 kernels + combine kernel
    int parts;        // partial sums per tile
    long long tiles_total;
    long long chunk_tiles;
    size_t partial_bytes, ws_bytes;
};

// 256 x 256 tiles: one position pair per thread QUAD (combine_quad.cuh) instead of per thread.  Opt-in
// (TFCFFT_COMBINE_QUAD=1): measured at parity with the one-thread item (profiles/r02_final_ab.txt: -3.5 % on luma, +1.5 %
// on per-channel RGB), so the one-thread item stays the product path.  Never with a generic-evaluation flag or one of the
// opt-in schedules that are built on the one-thread item.
inline bool combine_quad_enabled(long long p, unsigned flags) {
#ifdef TFCFFT_EMU_BUILD
    (void)p;
    (void)flags;
    return false;
#else
    static const bool on = getenv("TFCFFT_COMBINE_QUAD") != nullptr && getenv("TFCFFT_SUB_PIPE") == nullptr &&
                           !(getenv("TFCFFT_FINE_DEPS") != nullptr && getenv("TFCFFT_NO_PDL") == nullptr);
    return on && p == 256 && !(flags & (TFCFFT_LOG_MAGNITUDE | TFCFFT_FULL_SPECTRUM));
#endif
}
inline int split_parts(int p) {
    switch (p) {
        case 64: return Split<64>::PARTS;
        case 128: return Split<128>::PARTS;
        case 256: return Split<256>::PARTS;
        case 512: return Split<512>::PARTS;
        default: return 1;
    }
}

inline int validate_desc(const tfcfft_desc* d, Geometry* geo, bool allow_sub = true) {
    if (!d) return TFCFFT_ERR_NULL;
    if (d->struct_size != sizeof(tfcfft_desc)) return TFCFFT_ERR_STRUCT;
    if (elem_size(d->dtype) == 0) return TFCFFT_ERR_DTYPE;
    if (d->n == 0) return TFCFFT_ERR_EMPTY;
    if (d->n < 0 || d->n > (1 << 24)) return TFCFFT_ERR_SHAPE;
    if (d->c != 1 && d->c != 3) return TFCFFT_ERR_SHAPE;
    if (d->h != d->w || d->h <= 0 || d->grid <= 0 || d->h % d->grid) return TFCFFT_ERR_SHAPE;
    const long long p = d->h / d->grid;
    if (p != 16 && p != 32 && p != 64 && p != 128 && p != 256 && p != 512) return TFCFFT_ERR_SHAPE;
    const unsigned known = TFCFFT_CHANNELS_RGB | TFCFFT_NO_PHASE | TFCFFT_DIST_MSE | TFCFFT_PATCH_SUM |
                           TFCFFT_LOG_MAGNITUDE | TFCFFT_FULL_SPECTRUM | TFCFFT_QUANTIZE_U8 | TFCFFT_FORCE_SPLIT |
                           TFCFFT_FORCE_GENERIC | TFCFFT_USE_LINE | TFCFFT_USE_PAIR | TFCFFT_GRAD_ACCUMULATE | TFCFFT_USE_HALFLINE;
    if (d->flags & ~known) return TFCFFT_ERR_FLAGS;
    // the reference quantises to a single grey channel; a per-channel quantised variant does not exist
    if ((d->flags & TFCFFT_QUANTIZE_U8) && (d->flags & TFCFFT_CHANNELS_RGB) && d->c == 3) return TFCFFT_ERR_FLAGS;
    if ((d->flags & TFCFFT_FORCE_SPLIT) && p < 64) return TFCFFT_ERR_FLAGS;
    if (d->reserved != 0 || !(d->grad_scale_host == d->grad_scale_host)) return TFCFFT_ERR_FLAGS;
    for (int t = 0; t < 2; ++t) {
        const int64_t* st = t ? d->real_stride : d->fake_stride;
        if (st[3] != 1) return TFCFFT_ERR_STRIDE;
        for (int i = 0; i < 3; ++i)
            if (st[i] % 4 != 0 || st[i] < 0) return TFCFFT_ERR_STRIDE;
    }
    if (geo) {
        geo->p = (int)p;
        geo->luma3 = (d->c == 3) && !(d->flags & TFCFFT_CHANNELS_RGB);
        geo->cprime = (d->c == 3 && !geo->luma3) ? 3 : 1;
        static const bool no_d8 = getenv("TFCFFT_NO_D8") != nullptr;  // A/B switch: 512 x 512 tiles back on the split kernels
        geo->sub = allow_sub && (p == 128 || p == 256 || (p == 512 && !no_d8)) && !(d->flags & (TFCFFT_FORCE_SPLIT | TFCFFT_FORCE_GENERIC));
        geo->split = !geo->sub && ((p >= 256) || (d->flags & TFCFFT_FORCE_SPLIT));
        geo->parts = geo->sub ? (p == 512 ? kCombine8Parts : combine_quad_enabled(p, d->flags) ? kCombineQParts : kCombineParts)
                              : geo->split ? split_parts((int)p) : 1;
        geo->tiles_total = (long long)d->n * geo->cprime * d->grid * d->grid;
        geo->partial_bytes = align_up((size_t)geo->tiles_total * geo->parts * 2 * sizeof(float), 256);
        geo->chunk_tiles = 0;
        size_t z = 0;
        if (geo->split || geo->sub) {
            const size_t per_tile = (size_t)p * p * sizeof(float2);
            static const long long ws_mb = getenv("TFCFFT_WS_CHUNK_MB") ? atoll(getenv("TFCFFT_WS_CHUNK_MB")) : 0;  // A/B switch
            long long ct = (long long)((ws_mb > 0 ? (size_t)ws_mb << 20 : kWsChunkBytes) / per_tile);
            if (ct < 1) ct = 1;
            // (the launcher trims a chunk to whole waves of its launches from the device's occupancy: k_sub.cu)
            if (ct > kSchedMaxTiles) ct = kSchedMaxTiles;
            if (ct > geo->tiles_total) ct = geo->tiles_total;
            geo->chunk_tiles = ct;
            z = (size_t)ct * per_tile;
        }
        geo->ws_bytes = kWsHeader + geo->partial_bytes + z;
    }
    return TFCFFT_OK;
}

inline int check_grad_args(const tfcfft_desc* d, const void* grad) {
    if (!grad) return TFCFFT_OK;
    if ((d->flags & TFCFFT_QUANTIZE_U8) || d->dtype == TFCFFT_U8) return TFCFFT_ERR_NO_GRADIENT;
    if (d->grad_stride[3] != 1) return TFCFFT_ERR_STRIDE;
    for (int i = 0; i < 3; ++i)
        if (d->grad_stride[i] % 4 != 0 || d->grad_stride[i] < 0) return TFCFFT_ERR_STRIDE;
    return TFCFFT_OK;
}

inline int check_alignment(const tfcfft_desc* d, const void* fake, const void* real, const void* grad) {
    const uintptr_t a = 4 * elem_size(d->dtype);
    if ((uintptr_t)fake % a || (uintptr_t)real % a || (grad && (uintptr_t)grad % a)) return TFCFFT_ERR_ALIGNMENT;
    return TFCFFT_OK;
}

// Pillow's convert("L") coefficients as exact binary fractions (oracle/r1_differentiable.py LUMA_WEIGHTS)
constexpr float kLuma[3] = {19595.0f / 65536.0f, 38470.0f / 65536.0f, 7471.0f / 65536.0f};

inline Params make_params(const tfcfft_desc* d, const Geometry& g, const void* fake, const void* real, void* grad,
                          float* out, float* per_image, void* ws) {
    Params p{};
    p.fake = fake;
    p.real = real;
    p.grad = grad;
    for (int i = 0; i < 4; ++i) {
        p.fs[i] = d->fake_stride[i];
        p.rs[i] = d->real_stride[i];
        p.gs[i] = grad ? d->grad_stride[i] : 0;
    }
    p.n = (int)d->n; p.c = (int)d->c; p.h = (int)d->h; p.w = (int)d->w;
    p.grid = d->grid;
    p.p = g.p;
    p.cprime = g.cprime;
    p.tiles_per_image = g.cprime * d->grid * d->grid;
    p.tiles_total = (int)g.tiles_total;
    p.flags = d->flags;
    const float sc = (d->flags & TFCFFT_QUANTIZE_U8) ? 1.0f : d->input_scale;
    const float gsh = d->grad_scale_host == 0.f ? 1.f : d->grad_scale_host;
    for (int i = 0; i < 3; ++i) {
        p.lw[i] = g.luma3 ? kLuma[i] * sc : sc;
        p.gw[i] = p.lw[i] * gsh;
    }
    p.gscale_dev = d->grad_scale_dev;
    p.gscale_host = gsh;
    const double kbins = (d->flags & TFCFFT_FULL_SPECTRUM) ? (double)g.p * g.p : (double)g.p * (g.p / 2 + 1);
    const double red = (d->flags & TFCFFT_PATCH_SUM) ? (double)d->grid * d->grid : 1.0;
    p.norm = red / ((double)d->n * g.cprime * d->grid * d->grid * kbins);
    p.weight = d->weight;
    if (d->flags & TFCFFT_NO_PHASE) {
        p.sa = (float)((double)d->weight * p.norm);
        p.sp = 0.f;
    } else {
        p.sa = p.sp = (float)(0.5 * (double)d->weight * p.norm);
    }
    char* w = reinterpret_cast<char*>(ws);
    p.counter = reinterpret_cast<unsigned*>(w);
    p.partials = reinterpret_cast<float*>(w + kWsHeader);
    p.parts = g.parts;
    p.out = out;
    p.per_image = per_image;
    p.zws = (g.split || g.sub) ? reinterpret_cast<float2*>(w + kWsHeader + g.partial_bytes) : nullptr;
    p.sched = reinterpret_cast<unsigned*>(w + kSchedHeads);
    p.eq = reinterpret_cast<unsigned char*>(w + kEqFlags);
    p.sub_d = g.sub ? g.p / 64 : 0;
    p.tile_base = 0;
    p.chunk_tiles = (int)g.chunk_tiles;
    return p;
}

inline const char* status_string(int rc) {
    switch (rc) {
        case TFCFFT_OK: return "ok";
        case TFCFFT_ERR_NULL: return "tfcfft: null descriptor or required pointer";
        case TFCFFT_ERR_STRUCT: return "tfcfft: descriptor struct_size mismatch (ABI version skew)";
        case TFCFFT_ERR_DTYPE: return "tfcfft: unsupported dtype (f32, f16, bf16, u8)";
        case TFCFFT_ERR_SHAPE: return "tfcfft: unsupported shape (need C in {1,3}, H == W, H % grid == 0, patch side in {16,32,64,128,256,512})";
        case TFCFFT_ERR_STRIDE: return "tfcfft: unsupported strides (innermost must be 1, outer strides multiples of 4 elements)";
        case TFCFFT_ERR_ALIGNMENT: return "tfcfft: base pointer not aligned to 4 elements";
        case TFCFFT_ERR_FLAGS: return "tfcfft: unsupported flag combination";
        case TFCFFT_ERR_WORKSPACE: return "tfcfft: workspace null, misaligned or too small";
        case TFCFFT_ERR_NO_GRADIENT: return "tfcfft: no gradient exists for quantised (uint8) inputs";
        case TFCFFT_ERR_EMPTY: return "tfcfft: empty batch (N == 0)";
        default: return nullptr;
    }
}

// ---- regional 100 x 256 FFT loss ------------------------------------------------------------------------------
inline int validate_regional(const tfcfft_desc* d, Geometry* geo) {
    if (!d) return TFCFFT_ERR_NULL;
    if (d->struct_size != sizeof(tfcfft_desc)) return TFCFFT_ERR_STRUCT;
    if (elem_size(d->dtype) == 0) return TFCFFT_ERR_DTYPE;
    if (d->n == 0) return TFCFFT_ERR_EMPTY;
    if (d->n < 0 || d->n > (1 << 22)) return TFCFFT_ERR_SHAPE;
    if (d->c != 1 && d->c != 3) return TFCFFT_ERR_SHAPE;
    if (d->h != 256 || d->w != 256) return TFCFFT_ERR_SHAPE;  // the reference's bands: rows 0..99 and 100..199 of 256
    if (d->flags & ~(TFCFFT_CHANNELS_RGB | TFCFFT_NO_PHASE | TFCFFT_DIST_MSE | TFCFFT_QUANTIZE_U8 | TFCFFT_GRAD_ACCUMULATE))
        return TFCFFT_ERR_FLAGS;
    if ((d->flags & TFCFFT_QUANTIZE_U8) && (d->flags & TFCFFT_CHANNELS_RGB) && d->c == 3) return TFCFFT_ERR_FLAGS;
    for (int t = 0; t < 2; ++t) {
        const int64_t* st = t ? d->real_stride : d->fake_stride;
        if (st[3] != 1) return TFCFFT_ERR_STRIDE;
        for (int i = 0; i < 3; ++i)
            if (st[i] % 4 != 0 || st[i] < 0) return TFCFFT_ERR_STRIDE;
    }
    if (geo) {
        *geo = Geometry{};
        geo->p = 256;
        geo->luma3 = (d->c == 3) && !(d->flags & TFCFFT_CHANNELS_RGB);
        geo->cprime = (d->c == 3 && !geo->luma3) ? 3 : 1;
        geo->parts = 1;
        geo->tiles_total = (long long)d->n * geo->cprime * RegCfg::BANDS;
        geo->partial_bytes = align_up((size_t)geo->tiles_total * 2 * sizeof(float), 256);
        geo->ws_bytes = kWsHeader + geo->partial_bytes;
    }
    return TFCFFT_OK;
}

inline Params make_regional_params(const tfcfft_desc* d, const Geometry& g, const void* fake, const void* real, void* grad,
                                   float* out, float* per_image, void* ws) {
    tfcfft_desc d1 = *d;
    d1.grid = 1;
    Params p = make_params(&d1, g, fake, real, grad, out, per_image, ws);
    p.tiles_per_image = g.cprime * RegCfg::BANDS;
    p.tiles_total = (int)g.tiles_total;
    // nn.L1Loss per band = mean over N * C' * 100 * 129 bins; the two bands are SUMMED (withregion_FFT.py:398-399)
    p.norm = 1.0 / ((double)d->n * g.cprime * RegCfg::H * (RegCfg::W / 2 + 1));
    if (d->flags & TFCFFT_NO_PHASE) {
        p.sa = (float)((double)d->weight * p.norm);
        p.sp = 0.f;
    } else {
        p.sa = p.sp = (float)(0.5 * (double)d->weight * p.norm);
    }
    return p;
}

// ---- patch triplet loss ------------------------------------------------------------------------------------
constexpr int kTripletMaxBlocks = 4096;
constexpr size_t kTripletWsBytes = kWsHeader + (size_t)kTripletMaxBlocks * 2 * sizeof(float);

// The descriptor is validated like tfcfft_loss's, with these differences: grid in {1, 2, 4}, patch side a
// multiple of 16 up to 512 (no transform: no power-of-two rule), only TFCFFT_GRAD_ACCUMULATE as a flag.
inline int validate_triplet(const tfcfft_desc* d, const int32_t* negatives) {
    if (!d) return TFCFFT_ERR_NULL;
    if (d->struct_size != sizeof(tfcfft_desc)) return TFCFFT_ERR_STRUCT;
    if (elem_size(d->dtype) == 0) return TFCFFT_ERR_DTYPE;
    if (d->n == 0) return TFCFFT_ERR_EMPTY;
    if (d->n < 0 || d->n > (1 << 24)) return TFCFFT_ERR_SHAPE;
    if (d->c < 1 || d->c > 4) return TFCFFT_ERR_SHAPE;
    if (d->h != d->w || d->h <= 0) return TFCFFT_ERR_SHAPE;
    if (d->grid != 1 && d->grid != 2 && d->grid != 4) return TFCFFT_ERR_SHAPE;
    if (d->h % d->grid) return TFCFFT_ERR_SHAPE;
    const long long p = d->h / d->grid;
    if (p != 16 && p != 32 && p != 64 && p != 128 && p != 256 && p != 512) return TFCFFT_ERR_SHAPE;
    if (d->flags & ~TFCFFT_GRAD_ACCUMULATE) return TFCFFT_ERR_FLAGS;
    if (!negatives) return TFCFFT_ERR_NULL;
    for (int i = 0; i < d->grid * d->grid; ++i)
        if (negatives[i] < 0 || negatives[i] >= d->grid * d->grid) return TFCFFT_ERR_SHAPE;
    for (int t = 0; t < 2; ++t) {
        const int64_t* st = t ? d->real_stride : d->fake_stride;
        if (st[3] != 1) return TFCFFT_ERR_STRIDE;
        for (int i = 0; i < 3; ++i)
            if (st[i] % 4 != 0 || st[i] < 0) return TFCFFT_ERR_STRIDE;
    }
    return TFCFFT_OK;
}

// tfcfft_temperature_triplet: like validate_triplet, but grid is ignored (rows are whole image rows), flags may be
// QUANTIZE_U8 | GRAD_ACCUMULATE | TEMPS_POSITIVE, and the image side must be a power of two in 16..512
inline int validate_temperature(const tfcfft_desc* d, const int64_t* neg_stride) {
    if (!d) return TFCFFT_ERR_NULL;
    if (d->struct_size != sizeof(tfcfft_desc)) return TFCFFT_ERR_STRUCT;
    if (elem_size(d->dtype) == 0) return TFCFFT_ERR_DTYPE;
    if (d->n == 0) return TFCFFT_ERR_EMPTY;
    if (d->n < 0 || d->n > (1 << 24)) return TFCFFT_ERR_SHAPE;
    if (d->c < 1 || d->c > 4) return TFCFFT_ERR_SHAPE;
    if (d->h != d->w) return TFCFFT_ERR_SHAPE;
    const long long p = d->h;
    if (p != 16 && p != 32 && p != 64 && p != 128 && p != 256 && p != 512) return TFCFFT_ERR_SHAPE;
    if (d->flags & ~(TFCFFT_GRAD_ACCUMULATE | TFCFFT_QUANTIZE_U8 | TFCFFT_TEMPS_POSITIVE)) return TFCFFT_ERR_FLAGS;
    if (!neg_stride) return TFCFFT_ERR_NULL;
    for (int t = 0; t < 3; ++t) {
        const int64_t* st = t == 0 ? d->fake_stride : t == 1 ? d->real_stride : neg_stride;
        if (st[3] != 1) return TFCFFT_ERR_STRIDE;
        for (int i = 0; i < 3; ++i)
            if (st[i] % 4 != 0 || st[i] < 0) return TFCFFT_ERR_STRIDE;
    }
    return TFCFFT_OK;
}

inline TripletParams make_triplet_params(const tfcfft_desc* d, const void* fake, const void* real, const int32_t* negatives,
                                         float margin, float eps, float* out, void* grad, void* ws) {
    TripletParams t{};
    t.fake = fake;
    t.real = real;
    t.grad = grad;
    for (int i = 0; i < 4; ++i) {
        t.fs[i] = d->fake_stride[i];
        t.rs[i] = d->real_stride[i];
        t.gs[i] = grad ? d->grad_stride[i] : 0;
    }
    t.n = (int)d->n; t.c = (int)d->c; t.h = (int)d->h;
    t.grid = d->grid;
    t.p = (int)(d->h / d->grid);
    auto lg = [](long long v) { int r = 0; while ((1LL << r) < v) ++r; return r; };
    t.lg_g = lg(d->grid);
    t.lg_h = lg(d->h);
    t.lg_p = lg(t.p);
    for (int i = 0; i < 16; ++i) t.neg[i] = i < d->grid * d->grid ? negatives[i] : 0;
    t.margin = margin;
    t.eps = eps;
    t.weight = d->weight;
    t.rows = (long long)d->n * d->c * d->h * d->grid;
    t.coef = (float)((double)d->weight / (double)t.rows);
    t.accumulate = (d->flags & TFCFFT_GRAD_ACCUMULATE) ? 1 : 0;
    t.counter = reinterpret_cast<unsigned*>(ws);
    t.partials = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + kWsHeader);
    t.out = out;
    return t;
}

inline TripletParams make_temperature_params(const tfcfft_desc* d, const void* fake, const void* positive, const void* negative,
                                             const int64_t* neg_stride, const float* lut, float margin, float eps, float* out,
                                             void* grad, void* ws) {
    tfcfft_desc g1 = *d;
    g1.grid = 1;
    g1.flags &= TFCFFT_GRAD_ACCUMULATE;
    const int32_t self = 0;
    TripletParams t = make_triplet_params(&g1, fake, positive, &self, margin, eps, out, grad, ws);
    t.c = 1;  // the red channel only (datasets_temp.py:32)
    t.rows = (long long)d->n * d->h;
    t.neg_src = negative;
    for (int i = 0; i < 4; ++i) t.ns[i] = neg_stride[i];
    t.mode = (d->flags & TFCFFT_QUANTIZE_U8) ? 1 : 2;
    t.pos_f32 = (d->flags & TFCFFT_TEMPS_POSITIVE) ? 1 : 0;
    for (int i = 0; i < 256; ++i) t.lut[i] = lut[i];
    // differentiable variant: the table's end-to-end linear law on x' = input_scale * x (x' in 0..255)
    t.lin_b = (lut[255] - lut[0]) / 255.0f * d->input_scale;
    t.lin_a = lut[0];
    t.coef = (float)((double)d->weight / (double)t.rows) * (t.mode == 2 ? t.lin_b : 0.f);
    return t;
}

}  // namespace tfcfft
