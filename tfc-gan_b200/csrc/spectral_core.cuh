// spectral_core.cuh -- the loss algebra of the FFT-loss path, shared by the sm_100a kernels
// (spectral_kernels.cu) and by the serial CPU emulation (emu.cu).
//
// What is computed (reference: TFC-GAN-FFT/TFCGAN_multigpu_patchFFT_16P.py:271-375, restated
// differentiably as R1 in oracle/r1_differentiable.py):
//   per tile:  z = f + i r  (f, r = luma or single channel of fake / real, scaled)
//              Z = FFT2(z);  2F(k) = Z(k) + conj Z(-k);  2R(k) = -i (Z(k) - conj Z(-k))
//              on the half plane kc in [0, P/2]:  amp = |F|, |R| (optionally log), pha = atan2
//              distance L1 / squared, accumulated per tile;
//              spectral gradient G(k) = gA F/|F| (or F/|F|^2) + gP iF/|F|^2, zero-extended;
//              grad tile = Re(unnormalised inverse FFT2(G)).
#pragma once
#include <cmath>
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "../../include/tfcfft.h"
#include "fft_core.cuh"

namespace tfcfft {

struct Params {
    const void* fake;
    const void* real;
    void* grad;            // nullptr: forward only
    long long fs[4], rs[4], gs[4];  // element strides (N, C, H, W); W stride is 1
    int n, c, h, w, grid, p;
    int cprime;            // spectra per tile position: 1 (luma / single channel) or 3 (rgb)
    int tiles_per_image;   // cprime * grid * grid
    int tiles_total;       // n * tiles_per_image
    unsigned flags;
    float lw[3];           // load weights: luma coefficients * input_scale (or input_scale)
    float gw[3];           // gradient weights on the way out (same numbers)
    float sa, sp;          // gradient scale of the amplitude / phase term
    double norm;           // red / (N * C' * g^2 * K)
    float weight;
    float* partials;       // [tiles_total * parts][2] (amp, pha) raw sums
    int parts;             // partial sums per tile (1 resident, #column-group pairs split)
    unsigned* counter;     // self-resetting ticket for the last-block finalise
    int fine_deps;         // sub-tile launches: tile-granular dependencies through `sched` instead of whole-grid waits
    int defer_finish;      // sub-tile path with a gradient: the combine launches only store their partial sums; they are
                           // summed by the CTAs appended to the last inverse launch (`fin_ctas` of them, the first one
                           // works) or, on the two-lane schedule, by a one-CTA launch after the join
    int fin_ctas;
    int lookahead;         // sub-tile forward launches: L2 look-ahead bits (1: the CTA's next unit once its own loads have
                           // landed, 2: its first unit before the whole-grid wait)
    unsigned* sched;       // pipelined sub-tile kernel: queue heads, exit ticket, per-tile done counters (workspace header)
    unsigned char* eq;     // sub-tile path: "fake == real" flag per forward load unit of the chunk (nullptr: not tracked)
    float* out;            // [8]: loss, amp, pha, non-finite flag, gradient scale applied, 3 reserved
    float* per_image;      // [N][2] or nullptr
    float2* zws;           // split path: spectrum workspace [chunk_tiles][P][P]
    int tile_base;         // split path: first tile of this chunk
    int chunk_tiles;
    long long* trace;      // debug: per-CTA stage timestamps (tfcfft_debug_trace), normally nullptr
    // sub-tile path (P = 128 / 256): the tile is decimated into D x D interleaved 64 x 64 sub-images
    int sub_d;             // 0 / 1: not used; 2, 4 or 8
    int chunk_now;         // tiles in the chunk being processed by this launch
    // spectra materialisation (fft_components / make_spectra): grid == 1, tile = n * C' + ch
    int spec_mode;         // 0 loss, 1 emit amp / phase of both inputs, 2 backward from d/d(amp, phase)
    int spec_shift;        // outputs (and incoming gradients) in np.fft.fftshift order over both axes
    float* spec_out[4];    // emit: amp(fake), pha(fake), amp(real), pha(real); [tiles][P][W]; may be null
    const float* spec_gin[2];  // backward: d loss / d amp, d loss / d pha of `fake`; may be null
    // gradient epilogue: d loss / d fake is multiplied by *gscale_dev (GradScaler's device scale; may be null) on the
    // way out -- gw[] already holds the host-side factors -- and added to the buffer with TFCFFT_GRAD_ACCUMULATE
    const float* gscale_dev;
    float gscale_host;     // already folded into gw[]; reported in out[4] together with *gscale_dev
    // loader quadrants (tfcfft_loss_quads): `real` given as 4 separate [N,C,P,P] tensors (grid == 2); else all null
    const void* real_q[4];
};

// ---------------------------------------------------------------------------------------------
// element IO per dtype: 4 consecutive pixels, quantisation (torchvision to_pil_image semantics)
// ---------------------------------------------------------------------------------------------
// Sub-tile path (sub_tile.cuh), launch 2: work items per tile and how they are cut into CTAs / partial sums
constexpr int kCombineItems = 64 * 32 + 2 * 33;
#ifndef TFCFFT_COMBINE_THREADS
#define TFCFFT_COMBINE_THREADS 128
#endif
constexpr int kCombineThreads = TFCFFT_COMBINE_THREADS;
#ifndef TFCFFT_COMBINE_REP
#define TFCFFT_COMBINE_REP 2
#endif
constexpr int kCombineRep = TFCFFT_COMBINE_REP;  // items per thread (a rolled loop)
constexpr int kCombineItemsPerPart = kCombineThreads * kCombineRep;
constexpr int kCombineParts = (kCombineItems + kCombineItemsPerPart - 1) / kCombineItemsPerPart;  // CTAs (= partial sums) per tile
// 256 x 256 tiles on the quad combine (combine_quad.cuh): a CTA walks kCombineQRows rows of the 64 x 64 position grid;
// one more CTA per tile runs the self-conjugate columns
#ifndef TFC_CQ_RPC
#define TFC_CQ_RPC 4
#endif
constexpr int kCombineQRows = TFC_CQ_RPC;
constexpr int kCombineQParts = 64 / kCombineQRows + 1;
constexpr int kCombine8Parts = 64;  // 512 x 512 tiles (combine8.cuh): one CTA per row of the 64 x 64 position grid

template <typename T> struct IO;

// Pixel tensors are read once and gradients written once per call: mark them streaming (ld/st.global.cs) so they
// do not displace the L2-resident spectrum workspace of the multi-launch paths.
#ifndef TFCFFT_STREAM_HINTS
#define TFCFFT_STREAM_HINTS 1
#endif
template <typename V>
TFC_HD V ld_stream(const void* p) {
#if defined(__CUDA_ARCH__) && TFCFFT_STREAM_HINTS
    return __ldcs(reinterpret_cast<const V*>(p));
#else
    return *reinterpret_cast<const V*>(p);
#endif
}
template <typename V>
TFC_HD void st_stream(void* p, V v) {
#if defined(__CUDA_ARCH__) && TFCFFT_STREAM_HINTS
    __stcs(reinterpret_cast<V*>(p), v);
#else
    *reinterpret_cast<V*>(p) = v;
#endif
}

template <> struct IO<float> {
    TFC_HD static void load4(const float* p, float* v) {
        const float4 t = ld_stream<float4>(p);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
    TFC_HD static void store4(float* p, const float* v) { st_stream<float4>(p, make_float4(v[0], v[1], v[2], v[3])); }
    TFC_HD static void store1(float* p, float v) { *p = v; }
    TFC_HD static void load2(const float* p, float* v) {
        const float2 t = ld_stream<float2>(p);
        v[0] = t.x; v[1] = t.y;
    }
    TFC_HD static void store2(float* p, float a, float b) { st_stream<float2>(p, make_float2(a, b)); }
    // (x * 255) in fp32, truncated toward zero, wrapped mod 256
    TFC_HD static int quant(float x) {
#ifdef __CUDA_ARCH__
        return __float2int_rz(__fmul_rn(x, 255.0f)) & 0xFF;
#else
        volatile float t = x * 255.0f;
        return ((int)t) & 0xFF;
#endif
    }
};

template <> struct IO<__half> {
    TFC_HD static void load4(const __half* p, float* v) {
        const uint2 t = ld_stream<uint2>(p);
        const __half2 a = *reinterpret_cast<const __half2*>(&t.x);
        const __half2 b = *reinterpret_cast<const __half2*>(&t.y);
        v[0] = __low2float(a); v[1] = __high2float(a); v[2] = __low2float(b); v[3] = __high2float(b);
    }
    TFC_HD static void store4(__half* p, const float* v) {
        const __half2 a = __floats2half2_rn(v[0], v[1]);
        const __half2 b = __floats2half2_rn(v[2], v[3]);
        uint2 t;
        t.x = *reinterpret_cast<const unsigned*>(&a);
        t.y = *reinterpret_cast<const unsigned*>(&b);
        st_stream<uint2>(p, t);
    }
    TFC_HD static void store1(__half* p, float v) { *p = __float2half_rn(v); }
    TFC_HD static void load2(const __half* p, float* v) {
        const unsigned u = ld_stream<unsigned>(p);
        const __half2 t = *reinterpret_cast<const __half2*>(&u);
        v[0] = __low2float(t); v[1] = __high2float(t);
    }
    TFC_HD static void store2(__half* p, float a, float b) {
        const __half2 h = __floats2half2_rn(a, b);
        st_stream<unsigned>(p, *reinterpret_cast<const unsigned*>(&h));
    }
    // fp16 * 255 rounded to fp16 (the exact fp32 product rounded once == the fp16 product)
    TFC_HD static int quant(float x) {
        const float t = __half2float(__float2half_rn(x * 255.0f));
        return ((int)t) & 0xFF;
    }
};

template <> struct IO<__nv_bfloat16> {
    TFC_HD static void load4(const __nv_bfloat16* p, float* v) {
        const uint2 t = ld_stream<uint2>(p);
        // bf16 -> fp32 is a 16-bit shift
        v[0] = __uint_as_float_hd(t.x << 16); v[1] = __uint_as_float_hd(t.x & 0xFFFF0000u);
        v[2] = __uint_as_float_hd(t.y << 16); v[3] = __uint_as_float_hd(t.y & 0xFFFF0000u);
    }
    TFC_HD static void store4(__nv_bfloat16* p, const float* v) {
        __nv_bfloat16 h[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = __float2bfloat16_rn(v[i]);
        st_stream<uint2>(p, *reinterpret_cast<const uint2*>(h));
    }
    TFC_HD static void store1(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
    TFC_HD static void load2(const __nv_bfloat16* p, float* v) {
        const unsigned t = ld_stream<unsigned>(p);
        v[0] = __uint_as_float_hd(t << 16); v[1] = __uint_as_float_hd(t & 0xFFFF0000u);
    }
    TFC_HD static void store2(__nv_bfloat16* p, float a, float b) {
        p[0] = __float2bfloat16_rn(a);
        p[1] = __float2bfloat16_rn(b);
    }
    TFC_HD static int quant(float x) { return IO<float>::quant(x); }  // NumPy has no bf16: fp32 rule
    TFC_HD static float __uint_as_float_hd(unsigned u) {
        float f;
        memcpy(&f, &u, 4);
        return f;
    }
};

template <> struct IO<uint8_t> {
    TFC_HD static void load4(const uint8_t* p, float* v) {
        const uchar4 t = *reinterpret_cast<const uchar4*>(p);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
    TFC_HD static void store4(uint8_t*, const float*) {}  // no gradient for integer inputs
    TFC_HD static void store1(uint8_t*, float) {}
    TFC_HD static void load2(const uint8_t* p, float* v) { v[0] = p[0]; v[1] = p[1]; }
    TFC_HD static void store2(uint8_t*, float, float) {}
    TFC_HD static int quant(float x) { return (int)x; }   // already an 8-bit code
};

// Reads pixels (y, x..x+3) of one tile from `base` (pointing at channel 0 of the tile origin) and
// returns the 4 transform inputs: luma of 3 channels (LUMA3) or the single channel, scaled.
template <typename T, bool LUMA3>
TFC_HD void load_px4(const Params& prm, const T* base, const long long* st, int y, int x, float* out) {
    const T* p0 = base + (long long)y * st[2] + x;
    const bool quant = (prm.flags & TFCFFT_QUANTIZE_U8) != 0;
    if constexpr (LUMA3) {
        float r[4], g[4], b[4];
        IO<T>::load4(p0, r);
        IO<T>::load4(p0 + st[1], g);
        IO<T>::load4(p0 + 2 * st[1], b);
        if (quant) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int l = (19595 * IO<T>::quant(r[i]) + 38470 * IO<T>::quant(g[i]) + 7471 * IO<T>::quant(b[i]) + 0x8000) >> 16;
                out[i] = (float)l;
            }
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) out[i] = fmaf(prm.lw[2], b[i], fmaf(prm.lw[1], g[i], prm.lw[0] * r[i]));
        }
    } else {
        float v[4];
        IO<T>::load4(p0, v);
#pragma unroll
        for (int i = 0; i < 4; ++i) out[i] = quant ? (float)IO<T>::quant(v[i]) : prm.lw[0] * v[i];
    }
}

struct TileCoord {
    int n, ch, py, px;
};
TFC_HD TileCoord decode_tile(const Params& prm, int tile) {
    TileCoord t;
    t.n = tile / prm.tiles_per_image;
    const int r = tile % prm.tiles_per_image;
    const int gg = prm.grid * prm.grid;
    t.ch = r / gg;
    const int pi = r % gg;
    t.py = pi / prm.grid;  // row-major tiles: B1..B4 is the top row (patchFFT_16P.py:234-251)
    t.px = pi % prm.grid;
    return t;
}
template <typename T>
TFC_HD const T* tile_ptr(const void* base, const long long* st, const TileCoord& t, int p) {
    return reinterpret_cast<const T*>(base) + t.n * st[0] + t.ch * st[1] + (long long)t.py * p * st[2] + (long long)t.px * p;
}

// `real` tile: a window of the full tensor, or -- loader quadrants B1..B4 = TL, TR, BL, BR (datasets_temp.py:76-118,
// fft_loss(fake_B, B1, B2, B3, B4) at TFCGAN_multigpu_patchFFT_experiment.py:317-339) -- a whole separate tensor
template <typename T>
TFC_HD const T* real_tile_ptr(const Params& prm, const TileCoord& t, int p) {
    if (prm.real_q[0] == nullptr) return tile_ptr<T>(prm.real, prm.rs, t, p);
    return reinterpret_cast<const T*>(prm.real_q[t.py * prm.grid + t.px]) + t.n * prm.rs[0] + t.ch * prm.rs[1];
}

// Gradient epilogue shared by every kernel: per-channel weights on the way out (luma coefficient x input_scale x
// host gradient scale, times the optional device scalar) and plain or accumulating 4-pixel stores.
struct GradOut {
    float w[3];
    bool acc;
};
TFC_HD GradOut grad_out(const Params& prm) {
    GradOut g;
    float sc = 1.f;
    if (prm.gscale_dev != nullptr) {
#ifdef __CUDA_ARCH__
        sc = __ldg(prm.gscale_dev);
#else
        sc = *prm.gscale_dev;
#endif
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) g.w[c] = prm.gw[c] * sc;
    g.acc = (prm.flags & TFCFFT_GRAD_ACCUMULATE) != 0;
    return g;
}
template <typename T>
TFC_HD void grad_store4(const GradOut& go, T* p, float* v) {
    if (go.acc) {
        float o[4];
        IO<T>::load4(p, o);
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] += o[i];
    }
    IO<T>::store4(p, v);
}
template <typename T>
TFC_HD void grad_store2(const GradOut& go, T* p, float a, float b) {
    if (go.acc) {
        float o[2];
        IO<T>::load2(p, o);
        a += o[0];
        b += o[1];
    }
    IO<T>::store2(p, a, b);
}

// Packs rows [row0, row0+nrows) of a tile into s[(r-row0)*ld + x] = (f, r).
template <int P, typename T, bool LUMA3, class Ctx>
TFC_HD void load_rows(const Ctx& ctx, const Params& prm, const TileCoord& tc, int row0, int nrows, float2* s, int ld) {
    const T* fb = tile_ptr<T>(prm.fake, prm.fs, tc, P);
    const T* rb = real_tile_ptr<T>(prm, tc, P);
    constexpr int XV = P / 4;
    for (int it = ctx.tid; it < nrows * XV; it += ctx.nthreads) {
        const int x = (it % XV) * 4, y = it / XV;
        float f[4], r[4];
        load_px4<T, LUMA3>(prm, fb, prm.fs, row0 + y, x, f);
        load_px4<T, LUMA3>(prm, rb, prm.rs, row0 + y, x, r);
        float2* d = s + y * ld + x;
#pragma unroll
        for (int i = 0; i < 4; ++i) d[i] = make_float2(f[i], r[i]);
    }
}

// Writes the gradient rows: grad_c = gw[c] * Re(s).
template <int P, typename T, bool LUMA3, class Ctx>
TFC_HD void store_rows(const Ctx& ctx, const Params& prm, const TileCoord& tc, int row0, int nrows, const float2* s, int ld) {
    T* gb = const_cast<T*>(tile_ptr<T>(prm.grad, prm.gs, tc, P));
    constexpr int XV = P / 4;
    const GradOut go = grad_out(prm);
    for (int it = ctx.tid; it < nrows * XV; it += ctx.nthreads) {
        const int x = (it % XV) * 4, y = it / XV;
        const float2* d = s + y * ld + x;
        T* p0 = gb + (long long)(row0 + y) * prm.gs[2] + x;
        constexpr int NC = LUMA3 ? 3 : 1;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            float v[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) v[i] = go.w[c] * d[i].x;
            grad_store4<T>(go, p0 + c * prm.gs[1], v);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// per-bin loss and spectral gradient
// ---------------------------------------------------------------------------------------------
TFC_HD float sgnf(float d) { return (float)((d > 0.f) - (d < 0.f)); }

// zk = Z(k), zm = Z(-k).  Accumulates mult * distance into accA / accP and returns G(k).
TFC_HD float2 bin_eval(const Params& prm, float2 zk, float2 zm, float mult, float& accA, float& accP) {
    // 2F and 2R.  At self-conjugate bins zk == zm, so the imaginary parts are exactly +0 and a
    // negative real bin gives phase +pi like NumPy / torch (SURVEY.md §7 hard part 5).
    const float fx = zk.x + zm.x, fy = zk.y - zm.y;
    const float rx = zk.y + zm.y, ry = zm.x - zk.x;
    const float f2 = sqrtf(fx * fx + fy * fy), r2 = sqrtf(rx * rx + ry * ry);  // 2|F|, 2|R|
    const float finv = f2 > 0.f ? 1.0f / f2 : 0.f;
    const bool mse = (prm.flags & TFCFFT_DIST_MSE) != 0;
    const bool lg = (prm.flags & TFCFFT_LOG_MAGNITUDE) != 0;
    float va = 0.5f * f2, vb = 0.5f * r2;
    if (lg) {
        va = logf(va);
        vb = logf(vb);
    }
    const float da = va - vb;
    float ga;
    if (mse) {
        accA += mult * da * da;
        ga = 2.f * da;
    } else {
        accA += mult * fabsf(da);
        ga = sgnf(da);
    }
    // d|F|/dF = F/|F| = F2/|F2|;  d log|F| / dF = F/|F|^2 = 2 F2/|F2|^2
    float ca = prm.sa * mult * ga * finv;
    if (lg) ca *= 2.f * finv;
    float2 g = make_float2(ca * fx, ca * fy);
    if (!(prm.flags & TFCFFT_NO_PHASE)) {
        const float dp = atan2f(fy, fx) - atan2f(ry, rx);
        float gp;
        if (mse) {
            accP += mult * dp * dp;
            gp = 2.f * dp;
        } else {
            accP += mult * fabsf(dp);
            gp = sgnf(dp);
        }
        // d angle(F)/dF = iF/|F|^2 = 2 i F2/|F2|^2
        const float cp = prm.sp * mult * gp * 2.f * finv * finv;
        g.x -= cp * fy;
        g.y += cp * fx;
    }
    return g;
}

// Column map of the resident tile: local column == column position.
template <int P>
struct TileCols {
    TFC_HD int freq(int cl) const { return freq_of_pos<P>(cl); }
    TFC_HD int partner(int cl) const { return neg_pos<P>(cl); }
};

// Column map of a split-path slab holding one or two groups of GS consecutive positions whose
// low frequency digits are k0 (local columns [0,GS)) and k1 = -k0 mod Q (local [GS, 2GS)).
template <int P>
struct SlabCols {
    static constexpr int GS = Plan<P>::R3 > 1 ? Plan<P>::R3 : Plan<P>::R2;
    static constexpr int Q = P / GS;
    int k0, k1;
    bool self;
    TFC_HD int freq(int cl) const { return (cl < GS ? k0 : k1) + Q * (cl % GS); }
    TFC_HD int partner(int cl) const {
        const int nk = (P - freq(cl)) & (P - 1);
        const int kl = nk / Q;  // nk % Q is the partner group's low digits by construction
        return ((self || cl >= GS) ? 0 : GS) + kl;
    }
    // position-space index of the group with low frequency digits kappa
    TFC_HD static int group_of(int kappa) {
        using Pl = Plan<P>;
        if constexpr (Pl::R3 > 1) return (kappa % Pl::R1) * Pl::R2 + kappa / Pl::R1;
        else return kappa;
    }
};

// ---- spectra materialisation (reference fft_components, patchFFT_16P.py:293-319; make_spectra :284-289) --
template <int P>
TFC_HD long long spec_index(const Params& prm, int tile, int kr, int kc) {
    const bool full = (prm.flags & TFCFFT_FULL_SPECTRUM) != 0;
    const int w = full ? P : P / 2 + 1;
    int r = kr, c = kc;
    if (prm.spec_shift) {  // np.fft.fftshift: out[(i + n/2) % n] = in[i]
        r = (kr + P / 2) % P;
        c = (kc + w / 2) % w;
    }
    return ((long long)tile * P + r) * w + c;
}

// zk = Z(k), zm = Z(-k) at half-plane bin k = (kr, kc): write amplitude / phase of F (and R).
template <int P>
TFC_HD void bin_emit(const Params& prm, int tile, int kr, int kc, float2 zk, float2 zm, bool mirror) {
    const float fx = zk.x + zm.x, fy = zk.y - zm.y;  // 2F
    const float rx = zk.y + zm.y, ry = zm.x - zk.x;  // 2R
    const bool lg = (prm.flags & TFCFFT_LOG_MAGNITUDE) != 0;
    float af = 0.5f * sqrtf(fx * fx + fy * fy), ar = 0.5f * sqrtf(rx * rx + ry * ry);
    if (lg) {
        af = logf(af);
        ar = logf(ar);
    }
    const float pf = atan2f(fy, fx), pr = atan2f(ry, rx);
    const long long i = spec_index<P>(prm, tile, kr, kc);
    if (prm.spec_out[0]) prm.spec_out[0][i] = af;
    if (prm.spec_out[1]) prm.spec_out[1][i] = pf;
    if (prm.spec_out[2]) prm.spec_out[2][i] = ar;
    if (prm.spec_out[3]) prm.spec_out[3][i] = pr;
    if (mirror) {  // full plane: bin -k has the conjugate value
        const long long m = spec_index<P>(prm, tile, (P - kr) & (P - 1), (P - kc) & (P - 1));
        if (prm.spec_out[0]) prm.spec_out[0][m] = af;
        if (prm.spec_out[1]) prm.spec_out[1][m] = atan2f(-fy, fx);
        if (prm.spec_out[2]) prm.spec_out[2][m] = ar;
        if (prm.spec_out[3]) prm.spec_out[3][m] = atan2f(-ry, rx);
    }
}

// Spectral gradient of bin k from incoming d/d amp, d/d phase (chain rule of abs / log / angle).
template <int P>
TFC_HD float2 bin_spec_bwd(const Params& prm, int tile, int kr, int kc, float2 zk, float2 zm, bool mirror) {
    const float fx = zk.x + zm.x, fy = zk.y - zm.y;  // 2F
    const float f2 = sqrtf(fx * fx + fy * fy);
    const float finv = f2 > 0.f ? 1.0f / f2 : 0.f;
    const long long i = spec_index<P>(prm, tile, kr, kc);
    float ga = prm.spec_gin[0] ? prm.spec_gin[0][i] : 0.f;
    float gp = prm.spec_gin[1] ? prm.spec_gin[1][i] : 0.f;
    if (mirror) {  // F(-k) = conj F(k): same amplitude, negated phase
        const long long m = spec_index<P>(prm, tile, (P - kr) & (P - 1), (P - kc) & (P - 1));
        if (prm.spec_gin[0]) ga += prm.spec_gin[0][m];
        if (prm.spec_gin[1]) gp -= prm.spec_gin[1][m];
    }
    float ca = ga * finv;                                     // d|F|/dF = F2/|F2|
    if (prm.flags & TFCFFT_LOG_MAGNITUDE) ca *= 2.f * finv;   // d log|F|/dF = 2 F2/|F2|^2
    const float cp = gp * 2.f * finv * finv;                  // d angle/dF = 2 i F2/|F2|^2
    return make_float2(ca * fx - cp * fy, ca * fy + cp * fx);
}

// Loss + spectral gradient over the spectrum held in s (rows = row positions, stride ld).
// Every half-plane bin is owned by exactly one item, which also owns its mirror position.
template <int P, class Cols, class Ctx>
TFC_HD void bin_pass(const Ctx& ctx, const Params& prm, int tile, float2* s, int ld, const Cols& cm, int ncols, float& accA,
                     float& accP) {
    const bool want_grad = prm.grad != nullptr;
    const bool full = (prm.flags & TFCFFT_FULL_SPECTRUM) != 0;
    const int mode = prm.spec_mode;
    for (int it = ctx.tid; it < P * ncols; it += ctx.nthreads) {
        const int cl = it % ncols, qr = it / ncols;
        const int kc = cm.freq(cl);
        if (kc > P / 2) continue;
        const int kr = freq_of_pos<P>(qr);
        const bool special = (kc == 0) || (kc == P / 2);  // self-conjugate columns
        if (special && kr > P / 2) continue;
        const int krm = (P - kr) & (P - 1);
        const int qrm = pos_of_freq<P>(krm);
        const int clm = cm.partner(cl);
        float2* pk = s + qr * ld + cl;
        float2* pm = s + qrm * ld + clm;
        const float2 zk = *pk, zm = *pm;
        if (mode == 1) {
            bin_emit<P>(prm, tile, kr, kc, zk, zm, full && !special);
            if (special && pm != pk) bin_emit<P>(prm, tile, krm, kc, zm, zk, false);
            continue;
        }
        const float mult = (full && !special) ? 2.f : 1.f;
        const float2 g = mode == 2 ? bin_spec_bwd<P>(prm, tile, kr, kc, zk, zm, full && !special)
                                   : bin_eval(prm, zk, zm, mult, accA, accP);
        if (!special) {
            if (want_grad) {
                *pk = g;
                *pm = make_float2(0.f, 0.f);
            }
        } else {
            if (pm != pk) {  // the mirrored row of the same column is a half-plane bin of its own
                const float2 g2 = mode == 2 ? bin_spec_bwd<P>(prm, tile, krm, kc, zm, zk, false)
                                            : bin_eval(prm, zm, zk, mult, accA, accP);
                if (want_grad) *pm = g2;
            }
            if (want_grad) *pk = g;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// resident-tile path (P <= 128): the whole complex tile lives in shared memory
// ---------------------------------------------------------------------------------------------
template <int P, typename T, bool LUMA3, class Ctx>
TFC_HD void tile_process(const Ctx& ctx, const Params& prm, int tile, float2* s, const float2* tw, float& accA, float& accP) {
    constexpr int LD = P + 1;  // odd row pitch (in complex elements): conflict-free row and column walks
    constexpr int LP = ilog2_c(P);
    const TileCoord tc = decode_tile(prm, tile);
    load_rows<P, T, LUMA3>(ctx, prm, tc, 0, P, s, LD);
    ctx.sync();
    fft_lines<P, false>(ctx, s, 1, LD, LP, tw);   // rows: thread-fast index = row
    fft_lines<P, false>(ctx, s, LD, 1, LP, tw);   // columns: thread-fast index = column
    bin_pass<P>(ctx, prm, tile, s, LD, TileCols<P>(), P, accA, accP);
    ctx.sync();
    if (prm.grad != nullptr) {
        fft_lines<P, true>(ctx, s, LD, 1, LP, tw);
        fft_lines<P, true>(ctx, s, 1, LD, LP, tw);
        store_rows<P, T, LUMA3>(ctx, prm, tc, 0, P, s, LD);
        ctx.sync();
    }
}

// ---------------------------------------------------------------------------------------------
// split path (P >= 256, or forced): rows -> workspace -> column-group pairs -> workspace -> rows
// ---------------------------------------------------------------------------------------------
template <int P>
struct Split {
    static constexpr int RS = (8192 / P) < P ? (8192 / P) : P;  // rows per row-slab (~64 KB)
    static constexpr int GS = SlabCols<P>::GS;
    static constexpr int Q = SlabCols<P>::Q;
    static constexpr int PARTS = Q / 2 + 1;                      // column-group pairs per tile
    static constexpr int ROW_SLABS = P / RS;
};

// forward rows of slab `slab` of local tile `lt` (global tile = tile_base + lt) -> workspace
template <int P, typename T, bool LUMA3, class Ctx>
TFC_HD void split_rows_fwd(const Ctx& ctx, const Params& prm, int lt, int slab, float2* s, const float2* tw) {
    constexpr int LD = P + 1, RS = Split<P>::RS;
    const TileCoord tc = decode_tile(prm, prm.tile_base + lt);
    load_rows<P, T, LUMA3>(ctx, prm, tc, slab * RS, RS, s, LD);
    ctx.sync();
    fft_lines<P, false>(ctx, s, 1, LD, ilog2_c(RS), tw);
    float2* z = prm.zws + ((long long)lt * P + slab * RS) * P;
    for (int it = ctx.tid; it < RS * P; it += ctx.nthreads) z[it] = s[(it / P) * LD + (it % P)];
    ctx.sync();
}

// columns of group pair `pair`: forward FFT, loss + spectral gradient, inverse FFT, back to workspace
template <int P, class Ctx>
TFC_HD void split_cols(const Ctx& ctx, const Params& prm, int lt, int pair, float2* s, const float2* tw, float& accA, float& accP) {
    using Sp = Split<P>;
    constexpr int GS = Sp::GS, Q = Sp::Q;
    SlabCols<P> cm;
    cm.k0 = pair;
    cm.k1 = (Q - pair) % Q;
    cm.self = (cm.k1 == cm.k0);
    const int ncols = cm.self ? GS : 2 * GS;
    const int ld = ncols + 1;
    const int c0 = SlabCols<P>::group_of(cm.k0) * GS, c1 = SlabCols<P>::group_of(cm.k1) * GS;
    float2* z = prm.zws + (long long)lt * P * P;
    for (int it = ctx.tid; it < P * ncols; it += ctx.nthreads) {
        const int cl = it % ncols, r = it / ncols;
        s[r * ld + cl] = z[(long long)r * P + (cl < GS ? c0 + cl : c1 + cl - GS)];
    }
    ctx.sync();
    fft_lines<P, false>(ctx, s, ld, 1, ilog2(ncols), tw);
    bin_pass<P>(ctx, prm, prm.tile_base + lt, s, ld, cm, ncols, accA, accP);
    ctx.sync();
    if (prm.grad != nullptr) {
        fft_lines<P, true>(ctx, s, ld, 1, ilog2(ncols), tw);
        for (int it = ctx.tid; it < P * ncols; it += ctx.nthreads) {
            const int cl = it % ncols, r = it / ncols;
            z[(long long)r * P + (cl < GS ? c0 + cl : c1 + cl - GS)] = s[r * ld + cl];
        }
        ctx.sync();
    }
}

// inverse rows of slab `slab` -> gradient
template <int P, typename T, bool LUMA3, class Ctx>
TFC_HD void split_rows_inv(const Ctx& ctx, const Params& prm, int lt, int slab, float2* s, const float2* tw) {
    constexpr int LD = P + 1, RS = Split<P>::RS;
    const TileCoord tc = decode_tile(prm, prm.tile_base + lt);
    const float2* z = prm.zws + ((long long)lt * P + slab * RS) * P;
    for (int it = ctx.tid; it < RS * P; it += ctx.nthreads) s[(it / P) * LD + (it % P)] = z[it];
    ctx.sync();
    fft_lines<P, true>(ctx, s, 1, LD, ilog2_c(RS), tw);
    store_rows<P, T, LUMA3>(ctx, prm, tc, slab * RS, RS, s, LD);
    ctx.sync();
}

// ---------------------------------------------------------------------------------------------
// final reduction: fixed order, double accumulation -> run-to-run bit-stable loss
// ---------------------------------------------------------------------------------------------
// Sum of image `img`'s partials in tile order.
TFC_HD void image_sums(const Params& prm, int img, double& a, double& p) {
    const float* q = prm.partials + (long long)img * prm.tiles_per_image * prm.parts * 2;
    a = 0.0;
    p = 0.0;
    for (int i = 0; i < prm.tiles_per_image * prm.parts; ++i) {
        a += (double)q[2 * i];
        p += (double)q[2 * i + 1];
    }
}
TFC_HD void write_outputs(const Params& prm, double suma, double sump) {
    const double amp = suma * prm.norm, pha = sump * prm.norm;
    const bool use_phase = !(prm.flags & TFCFFT_NO_PHASE);
    const double loss = use_phase ? (double)prm.weight * 0.5 * (amp + pha) : (double)prm.weight * amp;
    prm.out[0] = (float)loss;
    prm.out[1] = (float)amp;
    prm.out[2] = (float)pha;
    prm.out[3] = (loss - loss == 0.0) ? 0.f : 1.f;  // 1 when the loss is inf / nan
    // the factor folded into the gradient on top of d loss / d fake (tfcfft_grad_rescale's `applied` scalar)
    float sc = prm.gscale_host;
    if (prm.gscale_dev != nullptr) sc *= *prm.gscale_dev;
    prm.out[4] = sc;
}

// the packed / thread-per-line 64 x 64 engines implement the default loss modes only
TFC_HD bool pair_supported(const Params& prm) {
    return prm.p == 64 && prm.spec_mode == 0 && !(prm.flags & (TFCFFT_LOG_MAGNITUDE | TFCFFT_FULL_SPECTRUM | TFCFFT_FORCE_SPLIT | TFCFFT_FORCE_GENERIC));
}

}  // namespace tfcfft
