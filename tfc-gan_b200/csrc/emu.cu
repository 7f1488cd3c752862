// emu.cu -- serial CPU execution of the kernels' own arithmetic (libtfcfft_emu.so).
//
// TEST INFRASTRUCTURE.  This is NOT a fallback: the product package never loads it.  It exists so
// that the index arithmetic of the in-place digit-reversed FFT passes, the Hermitian un-mixing,
// the bin ownership rules and the reduction layout -- the parts of spectral_core.cuh that are easy
// to get subtly wrong -- can be checked against the oracle by the `-m "not gpu"` tests in a
// container without a GPU.  It runs the same __host__ __device__ templates the sm_100a kernels
// instantiate, with one serial "thread" (SerialCtx) instead of a thread block.
#include <cstdlib>
#include <vector>

#include "host_common.h"
#include "pair_tile.cuh"
#include "sub_tile.cuh"
#include "line_tile.cuh"

using namespace tfcfft;

namespace {

template <int P, typename T, bool LUMA3>
void run_resident(Params prm) {
    SerialCtx ctx;
    std::vector<float2> s((size_t)P * (P + 1)), tw(P);
    fill_twiddles<P>(ctx, tw.data());
    for (int tile = 0; tile < prm.tiles_total; ++tile) {
        float a = 0.f, p = 0.f;
        tile_process<P, T, LUMA3>(ctx, prm, tile, s.data(), tw.data(), a, p);
        prm.partials[2 * tile] = a;
        prm.partials[2 * tile + 1] = p;
    }
}

template <typename T, bool LUMA3>
void run_line(Params prm) {
    SerialCtx ctx;
    std::vector<float2> s((size_t)64 * LineCfg::LD);
    for (int tile = 0; tile < prm.tiles_total; ++tile) {
        float a = 0.f, p = 0.f;
        line_process<T, LUMA3>(ctx, prm, tile, s.data(), a, p);
        prm.partials[2 * tile] = a;
        prm.partials[2 * tile + 1] = p;
    }
}

template <int P, typename T, bool LUMA3>
void run_pair(Params prm) {
    if constexpr (P == 64) {
        SerialCtx ctx;
        std::vector<float4> s((size_t)P * PairCfg<P>::LD), tw(2 * P);
        fill_twiddles4<P>(ctx, tw.data());
        fill_row_twiddles4<P>(ctx, tw.data(), tw.data() + P);
        for (int ta = 0; ta < prm.tiles_total; ta += 2) {
            const bool b_valid = ta + 1 < prm.tiles_total;
            const int tb = b_valid ? ta + 1 : ta;
            float2 a = make_float2(0.f, 0.f), p = make_float2(0.f, 0.f);
            pair_process<P, T, LUMA3>(ctx, prm, ta, tb, b_valid, s.data(), tw.data(), a, p);
            prm.partials[2 * ta] = a.x;
            prm.partials[2 * ta + 1] = p.x;
            if (b_valid) {
                prm.partials[2 * tb] = a.y;
                prm.partials[2 * tb + 1] = p.y;
            }
        }
    }
}

template <typename T, bool LUMA3>
void run_sub(Params prm) {
    SerialCtx ctx;
    const int D = prm.sub_d, npp = D * D / 2;
    std::vector<float2> s((size_t)2 * 64 * SubCfg::LD);
    for (int base = 0; base < prm.tiles_total; base += prm.chunk_tiles) {
        prm.tile_base = base;
        prm.chunk_now = prm.tiles_total - base < prm.chunk_tiles ? prm.tiles_total - base : prm.chunk_tiles;
        if (D == 4) {  // the cluster variant's load (sub_fwd_load_quad): both halves, both column pairs, then the transforms
            std::vector<float2> s2((size_t)2 * 64 * SubCfg::LD);
            for (int w = 0; w < prm.chunk_now * 4; ++w) {
                const TileCoord tc = decode_tile(prm, base + (w >> 2));
                for (int half = 0; half < 2; ++half) sub_fwd_load_quad<T, LUMA3>(ctx, prm, tc, w & 3, half, s.data(), s2.data());
                for (int i = 0; i < 2; ++i) {
                    SubUnit su;
                    su.tile_local = w >> 2;
                    su.p = w & 3;
                    su.i = i;
                    su.plane = su.p * 2 + i;
                    float2* t = i ? s2.data() : s.data();
                    sub_fwd_rows(ctx, t);
                    sub_fwd_cols_store(ctx, prm, su, t);
                }
            }
        } else {
            for (int u = 0; u < prm.chunk_now * npp; ++u) sub_fwd_process<T, LUMA3>(ctx, prm, u, s.data());
        }
        for (int lt = 0; lt < prm.chunk_now; ++lt) {
            float2* ws_tile = sub_plane(prm, lt, 0);
            for (int part = 0; part < kCombineParts; ++part) {
                float a = 0.f, p = 0.f;
                for (int item = part * kCombineItemsPerPart; item < (part + 1) * kCombineItemsPerPart && item < kCombineItems; ++item) {
                    if (D == 2) combine_item<2>(prm, ws_tile, item, a, p);
                    else combine_item<4>(prm, ws_tile, item, a, p);
                }
                prm.partials[2 * ((size_t)(base + lt) * kCombineParts + part)] = a;
                prm.partials[2 * ((size_t)(base + lt) * kCombineParts + part) + 1] = p;
            }
        }
        if (prm.grad && D == 4) {  // the cluster variant's store (sub_inv_store_quad)
            std::vector<float2> t0((size_t)64 * SubCfg::LD), t1((size_t)64 * SubCfg::LD);
            for (int w = 0; w < prm.chunk_now * 4; ++w) {
                for (int i = 0; i < 2; ++i) {
                    SubUnit su;
                    su.tile_local = w >> 2;
                    su.p = w & 3;
                    su.i = i;
                    su.plane = su.p * 2 + i;
                    float2* t = i ? t1.data() : t0.data();
                    sub_inv_cols(ctx, prm, su, t);
                    sub_inv_rows(ctx, t);
                }
                const TileCoord tc = decode_tile(prm, base + (w >> 2));
                for (int half = 0; half < 2; ++half) sub_inv_store_quad<T, LUMA3>(ctx, prm, tc, w & 3, half, t0.data(), t1.data());
            }
        } else if (prm.grad) {
            for (int u = 0; u < prm.chunk_now * npp; ++u) sub_inv_process<T, LUMA3>(ctx, prm, u, s.data());
        }
    }
}

template <int P, typename T, bool LUMA3>
void run_split(Params prm) {
    if constexpr (P >= 64) {
        using Sp = Split<P>;
        SerialCtx ctx;
        std::vector<float2> s((size_t)P * (2 * Sp::GS + 1) + (size_t)Sp::RS * (P + 1)), tw(P);
        fill_twiddles<P>(ctx, tw.data());
        for (int base = 0; base < prm.tiles_total; base += prm.chunk_tiles) {
            prm.tile_base = base;
            const int nt = prm.tiles_total - base < prm.chunk_tiles ? prm.tiles_total - base : prm.chunk_tiles;
            for (int lt = 0; lt < nt; ++lt)
                for (int sl = 0; sl < Sp::ROW_SLABS; ++sl) split_rows_fwd<P, T, LUMA3>(ctx, prm, lt, sl, s.data(), tw.data());
            for (int lt = 0; lt < nt; ++lt)
                for (int pr = 0; pr < Sp::PARTS; ++pr) {
                    float a = 0.f, p = 0.f;
                    split_cols<P>(ctx, prm, lt, pr, s.data(), tw.data(), a, p);
                    prm.partials[2 * ((size_t)(base + lt) * Sp::PARTS + pr)] = a;
                    prm.partials[2 * ((size_t)(base + lt) * Sp::PARTS + pr) + 1] = p;
                }
            if (prm.grad)
                for (int lt = 0; lt < nt; ++lt)
                    for (int sl = 0; sl < Sp::ROW_SLABS; ++sl) split_rows_inv<P, T, LUMA3>(ctx, prm, lt, sl, s.data(), tw.data());
        }
    }
}

template <int P, typename T, bool LUMA3>
void run(const Params& prm, bool split) {
    if ((P == 128 || P == 256) && prm.sub_d > 1) run_sub<T, LUMA3>(prm);
    else if (split) run_split<P, T, LUMA3>(prm);
    else if (P == 64 && pair_supported(prm) && !(prm.flags & TFCFFT_USE_PAIR)) run_line<T, LUMA3>(prm);
    else if (P == 64 && pair_supported(prm)) run_pair<P, T, LUMA3>(prm);
    else if constexpr (P <= 128) run_resident<P, T, LUMA3>(prm);
}

template <int P, typename T>
void run_l(const Params& prm, bool split, bool luma3) {
    if (luma3) run<P, T, true>(prm, split);
    else run<P, T, false>(prm, split);
}

template <int P>
void run_t(const Params& prm, bool split, bool luma3, int dtype) {
    switch (dtype) {
        case TFCFFT_F32: run_l<P, float>(prm, split, luma3); break;
        case TFCFFT_F16: run_l<P, __half>(prm, split, luma3); break;
        case TFCFFT_BF16: run_l<P, __nv_bfloat16>(prm, split, luma3); break;
        case TFCFFT_U8: run_l<P, uint8_t>(prm, split, luma3); break;
    }
}

}  // namespace

static void run_all(Params& prm, const Geometry& g, int dtype) {
    switch (g.p) {
        case 16: run_t<16>(prm, g.split, g.luma3, dtype); break;
        case 32: run_t<32>(prm, g.split, g.luma3, dtype); break;
        case 64: run_t<64>(prm, g.split, g.luma3, dtype); break;
        case 128: run_t<128>(prm, g.split, g.luma3, dtype); break;
        case 256: run_t<256>(prm, g.split, g.luma3, dtype); break;
        case 512: run_t<512>(prm, g.split, g.luma3, dtype); break;
    }
}

// Same contract as tfcfft_loss, but every pointer is HOST memory and no workspace is passed in.
extern "C" int tfcfft_emulate(const tfcfft_desc* d, const void* fake, const void* real, float* out, float* per_image,
                              void* grad_fake) {
    Geometry g;
    int rc = validate_desc(d, &g);
    if (rc) return rc;
    if (!fake || !real || !out) return TFCFFT_ERR_NULL;
    if ((rc = check_grad_args(d, grad_fake))) return rc;
    if ((rc = check_alignment(d, fake, real, grad_fake))) return rc;
    std::vector<char> ws(g.ws_bytes, 0);
    Params prm = make_params(d, g, fake, real, grad_fake, out, per_image, ws.data());
    run_all(prm, g, d->dtype);
    double sa = 0.0, sp = 0.0;
    for (int img = 0; img < prm.n; ++img) {
        double a, p;
        image_sums(prm, img, a, p);
        if (per_image) {
            per_image[2 * img] = (float)(a * prm.norm * prm.n);
            per_image[2 * img + 1] = (float)(p * prm.norm * prm.n);
        }
        sa += a;
        sp += p;
    }
    write_outputs(prm, sa, sp);
    return TFCFFT_OK;
}

// Host-memory twins of tfcfft_spectra / tfcfft_spectra_bwd.
extern "C" int tfcfft_emulate_spectra(const tfcfft_desc* d, const void* x, const void* y, float* amp_x, float* pha_x,
                                      float* amp_y, float* pha_y, int fftshift) {
    Geometry g;
    int rc = validate_desc(d, &g, false);
    if (rc) return rc;
    if (d->grid != 1) return TFCFFT_ERR_SHAPE;
    if (!x) return TFCFFT_ERR_NULL;
    std::vector<char> ws(g.ws_bytes, 0);
    float out[4];
    Params prm = make_params(d, g, x, y ? y : x, nullptr, out, nullptr, ws.data());
    prm.spec_mode = 1;
    prm.spec_shift = fftshift != 0;
    prm.spec_out[0] = amp_x;
    prm.spec_out[1] = pha_x;
    prm.spec_out[2] = y ? amp_y : nullptr;
    prm.spec_out[3] = y ? pha_y : nullptr;
    run_all(prm, g, d->dtype);
    return TFCFFT_OK;
}

extern "C" int tfcfft_emulate_spectra_bwd(const tfcfft_desc* d, const void* x, const float* grad_amp, const float* grad_pha,
                                          void* grad_x, int fftshift) {
    Geometry g;
    int rc = validate_desc(d, &g, false);
    if (rc) return rc;
    if (d->grid != 1) return TFCFFT_ERR_SHAPE;
    if (!x || !grad_x) return TFCFFT_ERR_NULL;
    if ((rc = check_grad_args(d, grad_x))) return rc;
    std::vector<char> ws(g.ws_bytes, 0);
    float out[4];
    Params prm = make_params(d, g, x, x, grad_x, out, nullptr, ws.data());
    prm.spec_mode = 2;
    prm.spec_shift = fftshift != 0;
    prm.spec_gin[0] = grad_amp;
    prm.spec_gin[1] = grad_pha;
    run_all(prm, g, d->dtype);
    return TFCFFT_OK;
}

// Host-memory twin of tfcfft_patch_triplet: one serial "lane" per patch row.
template <typename T>
static void run_triplet(const TripletParams& tp) {
    double sum = 0.0, act = 0.0;
    for (long long row = 0; row < tp.rows; ++row) {
        float l = 0.f, a = 0.f;
        const SerialReduce red;
        switch (tp.p) {
            case 16: triplet_row<T, 4>(tp, row, 0, 1, red, l, a); break;
            case 32: triplet_row<T, 8>(tp, row, 0, 1, red, l, a); break;
            case 64: triplet_row<T, 16>(tp, row, 0, 1, red, l, a); break;
            case 128: triplet_row<T, 32>(tp, row, 0, 1, red, l, a); break;
            case 256: triplet_row<T, 64>(tp, row, 0, 1, red, l, a); break;
            case 512: triplet_row<T, 128>(tp, row, 0, 1, red, l, a); break;
        }
        sum += l;
        act += a;
    }
    triplet_outputs(tp, sum, act);
}

extern "C" int tfcfft_emulate_triplet(const tfcfft_desc* d, const void* fake, const void* real, const int32_t* negatives,
                                      float margin, float eps, float* out, void* grad_fake) {
    int rc = validate_triplet(d, negatives);
    if (rc) return rc;
    if (!fake || !real || !out) return TFCFFT_ERR_NULL;
    if ((rc = check_grad_args(d, grad_fake))) return rc;
    if ((rc = check_alignment(d, fake, real, grad_fake))) return rc;
    std::vector<char> ws(kTripletWsBytes, 0);
    const TripletParams tp = make_triplet_params(d, fake, real, negatives, margin, eps, out, grad_fake, ws.data());
    switch (d->dtype) {
        case TFCFFT_F32: run_triplet<float>(tp); break;
        case TFCFFT_F16: run_triplet<__half>(tp); break;
        case TFCFFT_BF16: run_triplet<__nv_bfloat16>(tp); break;
        case TFCFFT_U8: run_triplet<uint8_t>(tp); break;
    }
    return TFCFFT_OK;
}

extern "C" int tfcfft_emulate_temperature_triplet(const tfcfft_desc* d, const void* fake, const void* positive, const void* negative,
                                                  const int64_t* neg_stride, const float* lut, float margin, float eps, float* out,
                                                  void* grad_fake) {
    int rc = validate_temperature(d, neg_stride);
    if (rc) return rc;
    if (!fake || !positive || !negative || !lut || !out) return TFCFFT_ERR_NULL;
    if ((rc = check_grad_args(d, grad_fake))) return rc;
    std::vector<char> ws(kTripletWsBytes, 0);
    const TripletParams tp = make_temperature_params(d, fake, positive, negative, neg_stride, lut, margin, eps, out, grad_fake, ws.data());
    switch (d->dtype) {
        case TFCFFT_F32: run_triplet<float>(tp); break;
        case TFCFFT_F16: run_triplet<__half>(tp); break;
        case TFCFFT_BF16: run_triplet<__nv_bfloat16>(tp); break;
        case TFCFFT_U8: run_triplet<uint8_t>(tp); break;
    }
    return TFCFFT_OK;
}

template <typename T, bool LUMA3>
static void run_regional(Params prm) {
    SerialCtx ctx;
    std::vector<float2> s((size_t)RegCfg::H * RegCfg::LD), tw(RegCfg::W), w100(100);
    fill_twiddles<RegCfg::W>(ctx, tw.data());
    reg_fill_w100(ctx, w100.data());
    for (int unit = 0; unit < prm.tiles_total; ++unit) {
        float a = 0.f, p = 0.f;
        regional_process<T, LUMA3>(ctx, prm, unit, s.data(), tw.data(), w100.data(), a, p);
        prm.partials[2 * unit] = a;
        prm.partials[2 * unit + 1] = p;
    }
}

extern "C" int tfcfft_emulate_regional(const tfcfft_desc* d, const void* fake, const void* real, float* out, float* per_image,
                                       void* grad_fake) {
    Geometry g;
    int rc = validate_regional(d, &g);
    if (rc) return rc;
    if (!fake || !real || !out) return TFCFFT_ERR_NULL;
    if ((rc = check_grad_args(d, grad_fake))) return rc;
    std::vector<char> ws(g.ws_bytes, 0);
    Params prm = make_regional_params(d, g, fake, real, grad_fake, out, per_image, ws.data());
    switch (d->dtype) {
        case TFCFFT_F32: g.luma3 ? run_regional<float, true>(prm) : run_regional<float, false>(prm); break;
        case TFCFFT_F16: g.luma3 ? run_regional<__half, true>(prm) : run_regional<__half, false>(prm); break;
        case TFCFFT_BF16: g.luma3 ? run_regional<__nv_bfloat16, true>(prm) : run_regional<__nv_bfloat16, false>(prm); break;
        case TFCFFT_U8: g.luma3 ? run_regional<uint8_t, true>(prm) : run_regional<uint8_t, false>(prm); break;
    }
    double sa = 0.0, sp = 0.0;
    for (int img = 0; img < prm.n; ++img) {
        double a, p;
        image_sums(prm, img, a, p);
        if (per_image) {
            per_image[2 * img] = (float)(a * prm.norm * prm.n);
            per_image[2 * img + 1] = (float)(p * prm.norm * prm.n);
        }
        sa += a;
        sp += p;
    }
    write_outputs(prm, sa, sp);
    return TFCFFT_OK;
}

// Host twins of tfcfft_regional_spectra / _bwd (mode 1: amp / pha out; mode 2: grad_amp / grad_pha in, grad_x out).
extern "C" int tfcfft_emulate_regional_spectra(const tfcfft_desc* d, const void* x, float* amp, float* pha, const float* grad_amp,
                                               const float* grad_pha, void* grad_x, int fftshift) {
    Geometry g;
    int rc = validate_regional(d, &g);
    if (rc) return rc;
    if (!x) return TFCFFT_ERR_NULL;
    if ((rc = check_grad_args(d, grad_x))) return rc;
    std::vector<char> ws(g.ws_bytes, 0);
    float out[4];
    Params prm = make_regional_params(d, g, x, x, grad_x, out, nullptr, ws.data());
    prm.spec_mode = grad_x ? 2 : 1;
    prm.spec_shift = fftshift != 0;
    prm.spec_out[0] = amp;
    prm.spec_out[1] = pha;
    prm.spec_gin[0] = grad_amp;
    prm.spec_gin[1] = grad_pha;
    switch (d->dtype) {
        case TFCFFT_F32: g.luma3 ? run_regional<float, true>(prm) : run_regional<float, false>(prm); break;
        case TFCFFT_F16: g.luma3 ? run_regional<__half, true>(prm) : run_regional<__half, false>(prm); break;
        case TFCFFT_BF16: g.luma3 ? run_regional<__nv_bfloat16, true>(prm) : run_regional<__nv_bfloat16, false>(prm); break;
        case TFCFFT_U8: g.luma3 ? run_regional<uint8_t, true>(prm) : run_regional<uint8_t, false>(prm); break;
    }
    return TFCFFT_OK;
}
