// emu.cu -- serial CPU execution of the kernels' own arithmetic (libtfcfft_emu.so).
//
// TEST INFRASTRUCTURE.  This is NOT a fallback: the product package never loads it.  It exists so
// that the index arithmetic of the in-place digit-reversed FFT passes, the Hermitian un-mixing,
// the bin ownership rules and the reduction layout -- the parts of spectral_core.cuh that are easy
// to get subtly wrong -- can be checked against the oracle by the `-m "not gpu"` tests in a
// container without a GPU.  It runs the same __host__ __device__ templates the sm_100a kernels
// instantiate, with one serial "thread" (SerialCtx) instead of a thread block.
#define TFCFFT_EMU_BUILD 1  // serial emulation: the one-thread combine item (9 partial sums per 256 x 256 tile)
#include <cstdlib>
#include <vector>

#include "host_common.h"
#include "pair_tile.cuh"
#include "sub_tile.cuh"
#include "line_tile.cuh"

using namespace tfcfft;

void emu_run_f32(Params& prm, const Geometry& g);   // emu_fft.cu, one object per element type
void emu_run_f16(Params& prm, const Geometry& g);
void emu_run_bf16(Params& prm, const Geometry& g);
void emu_run_u8(Params& prm, const Geometry& g);

static void run_all(Params& prm, const Geometry& g, int dtype) {
    switch (dtype) {
        case TFCFFT_F32: emu_run_f32(prm, g); break;
        case TFCFFT_F16: emu_run_f16(prm, g); break;
        case TFCFFT_BF16: emu_run_bf16(prm, g); break;
        case TFCFFT_U8: emu_run_u8(prm, g); break;
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
    float out[8];
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
    float out[8];
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
    float out[8];
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
