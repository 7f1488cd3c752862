// tfcfft_api.cu -- the C-ABI entry points of libtfcfft.so (include/tfcfft.h).
//
// Host side only validates, builds the kernel parameter block and enqueues launches on the
// caller's stream.  No device allocation, no host synchronisation, no fallback of any kind.
// The kernels live in the k_*.cu translation units (launchers.h).
#include <atomic>
#include <cstdio>

#include "launchers.h"

using namespace tfcfft;

namespace tfcfft {
std::atomic<long long> g_launches{0};
}

namespace {

std::atomic<long long*> g_trace{nullptr};  // debug only (tfcfft_debug_trace)

// Which engine runs a validated descriptor
int dispatch(const Params& prm, const Geometry& g, int dtype, cudaStream_t st) {
    if (prm.sub_d > 1) return launch_sub_any(dtype, g.luma3, prm, st);
    if (g.split) return launch_split_any(g.p, dtype, g.luma3, prm, st);
    if (g.p == 64 && pair_supported(prm)) {
        // measured (profiles/): the thread-per-line engine wins on both luma and single-channel tiles; the packed
        // pair kernel stays selectable for A/B runs
        if (prm.flags & TFCFFT_USE_PAIR) return launch_pair_any(dtype, g.luma3, prm, st);
        return launch_line_any(dtype, g.luma3, prm, st);
    }
    if (g.p <= 128) return launch_resident_any(g.p, dtype, g.luma3, prm, st);
    return TFCFFT_ERR_SHAPE;
}

}  // namespace

extern "C" {

int tfcfft_version(void) { return TFCFFT_VERSION; }

const char* tfcfft_strerror(int rc) {
    if (rc > 0) return cudaGetErrorString((cudaError_t)rc);
    const char* s = status_string(rc);
    return s ? s : "tfcfft: unknown status";
}

int tfcfft_validate(const tfcfft_desc* d) { return validate_desc(d, nullptr); }

size_t tfcfft_workspace_bytes(const tfcfft_desc* d) {
    Geometry g;
    if (validate_desc(d, &g) != TFCFFT_OK) return 0;
    return g.ws_bytes;
}

size_t tfcfft_spectra_workspace_bytes(const tfcfft_desc* d) {
    Geometry g;
    if (validate_desc(d, &g, /*allow_sub=*/false) != TFCFFT_OK) return 0;
    return g.ws_bytes;
}

int tfcfft_workspace_init(void* workspace, size_t workspace_bytes, void* stream) {
    if (!workspace || workspace_bytes < kWsHeader || ((uintptr_t)workspace & 255)) return TFCFFT_ERR_WORKSPACE;
    cudaError_t e = cudaMemsetAsync(workspace, 0, kWsHeader, (cudaStream_t)stream);
    return e == cudaSuccess ? 0 : (int)e;
}

// shared argument checks of the two spectra entry points; the scalar outputs of the reduction tail land in
// the workspace header (nobody reads them)
static int spectra_common(const tfcfft_desc* d, const void* x, const void* y, void* grad, void* workspace,
                          size_t workspace_bytes, Geometry* g, Params* prm) {
    int rc = validate_desc(d, g, /*allow_sub=*/false);
    if (rc) return rc;
    if (d->grid != 1) return TFCFFT_ERR_SHAPE;
    if (!x) return TFCFFT_ERR_NULL;
    if ((rc = check_grad_args(d, grad))) return rc;
    if ((rc = check_alignment(d, x, y ? y : x, grad))) return rc;
    if (!workspace || workspace_bytes < g->ws_bytes || ((uintptr_t)workspace & 255)) return TFCFFT_ERR_WORKSPACE;
    float* scratch_out = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + 64);
    *prm = make_params(d, *g, x, y ? y : x, grad, scratch_out, nullptr, workspace);
    return 0;
}

int tfcfft_spectra(const tfcfft_desc* d, const void* x, const void* y, float* amp_x, float* pha_x, float* amp_y,
                   float* pha_y, int fftshift, void* workspace, size_t workspace_bytes, void* stream) {
    Geometry g;
    Params prm;
    if (int rc = spectra_common(d, x, y, nullptr, workspace, workspace_bytes, &g, &prm)) return rc;
    prm.spec_mode = 1;
    prm.spec_shift = fftshift != 0;
    prm.spec_out[0] = amp_x;
    prm.spec_out[1] = pha_x;
    prm.spec_out[2] = y ? amp_y : nullptr;
    prm.spec_out[3] = y ? pha_y : nullptr;
    return dispatch(prm, g, d->dtype, (cudaStream_t)stream);
}

int tfcfft_spectra_bwd(const tfcfft_desc* d, const void* x, const float* grad_amp, const float* grad_pha, void* grad_x,
                       int fftshift, void* workspace, size_t workspace_bytes, void* stream) {
    Geometry g;
    Params prm;
    if (!grad_x) return TFCFFT_ERR_NULL;
    if (int rc = spectra_common(d, x, nullptr, grad_x, workspace, workspace_bytes, &g, &prm)) return rc;
    prm.spec_mode = 2;
    prm.spec_shift = fftshift != 0;
    prm.spec_gin[0] = grad_amp;
    prm.spec_gin[1] = grad_pha;
    // the generic kernels scale the outgoing gradient by gw = (luma weight) * input_scale: exactly d x'/d x
    return dispatch(prm, g, d->dtype, (cudaStream_t)stream);
}

static int loss_common(const tfcfft_desc* d, const void* fake, const void* real, const void* const* quads, float* out,
                       float* per_image, void* grad_fake, void* workspace, size_t workspace_bytes, void* stream) {
    Geometry g;
    int rc = validate_desc(d, &g);
    if (rc) return rc;
    if (!fake || !real || !out) return TFCFFT_ERR_NULL;
    if ((rc = check_grad_args(d, grad_fake))) return rc;
    if ((rc = check_alignment(d, fake, real, grad_fake))) return rc;
    if (!workspace || workspace_bytes < g.ws_bytes || ((uintptr_t)workspace & 255)) return TFCFFT_ERR_WORKSPACE;
    Params prm = make_params(d, g, fake, real, grad_fake, out, per_image, workspace);
    if (quads) {
        if (d->grid != 2) return TFCFFT_ERR_SHAPE;
        for (int i = 0; i < 4; ++i) {
            if (!quads[i]) return TFCFFT_ERR_NULL;
            if ((uintptr_t)quads[i] % (4 * elem_size(d->dtype))) return TFCFFT_ERR_ALIGNMENT;
            prm.real_q[i] = quads[i];
        }
    }
    prm.trace = g_trace.load();
    cudaStream_t st = (cudaStream_t)stream;
    return dispatch(prm, g, d->dtype, st);
}

int tfcfft_loss(const tfcfft_desc* d, const void* fake, const void* real, float* out, float* per_image, void* grad_fake,
                void* workspace, size_t workspace_bytes, void* stream) {
    return loss_common(d, fake, real, nullptr, out, per_image, grad_fake, workspace, workspace_bytes, stream);
}

int tfcfft_loss_quads(const tfcfft_desc* d, const void* fake, const void* const real_quadrants[4], float* out, float* per_image,
                      void* grad_fake, void* workspace, size_t workspace_bytes, void* stream) {
    if (!real_quadrants) return TFCFFT_ERR_NULL;
    return loss_common(d, fake, real_quadrants[0], real_quadrants, out, per_image, grad_fake, workspace, workspace_bytes, stream);
}

size_t tfcfft_regional_workspace_bytes(const tfcfft_desc* d) {
    Geometry g;
    if (validate_regional(d, &g) != TFCFFT_OK) return 0;
    return g.ws_bytes;
}

int tfcfft_regional_loss(const tfcfft_desc* d, const void* fake, const void* real, float* out, float* per_image, void* grad_fake,
                         void* workspace, size_t workspace_bytes, void* stream) {
    Geometry g;
    int rc = validate_regional(d, &g);
    if (rc) return rc;
    if (!fake || !real || !out) return TFCFFT_ERR_NULL;
    if ((rc = check_grad_args(d, grad_fake))) return rc;
    if ((rc = check_alignment(d, fake, real, grad_fake))) return rc;
    if (!workspace || workspace_bytes < g.ws_bytes || ((uintptr_t)workspace & 255)) return TFCFFT_ERR_WORKSPACE;
    const Params prm = make_regional_params(d, g, fake, real, grad_fake, out, per_image, workspace);
    cudaStream_t st = (cudaStream_t)stream;
    return launch_regional_any(d->dtype, g.luma3, prm, st);
}

static int regional_spectra_common(const tfcfft_desc* d, const void* x, void* grad, void* workspace, size_t workspace_bytes,
                                   int mode, int fftshift, float* const* outs, const float* const* gins, cudaStream_t st) {
    Geometry g;
    int rc = validate_regional(d, &g);
    if (rc) return rc;
    if (d->flags & (TFCFFT_NO_PHASE | TFCFFT_DIST_MSE)) return TFCFFT_ERR_FLAGS;
    if (!x) return TFCFFT_ERR_NULL;
    if ((rc = check_grad_args(d, grad))) return rc;
    if ((rc = check_alignment(d, x, x, grad))) return rc;
    if (!workspace || workspace_bytes < g.ws_bytes || ((uintptr_t)workspace & 255)) return TFCFFT_ERR_WORKSPACE;
    float* scratch_out = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + 64);
    Params prm = make_regional_params(d, g, x, x, grad, scratch_out, nullptr, workspace);
    prm.spec_mode = mode;
    prm.spec_shift = fftshift != 0;
    if (outs) {
        prm.spec_out[0] = outs[0];
        prm.spec_out[1] = outs[1];
    }
    if (gins) {
        prm.spec_gin[0] = gins[0];
        prm.spec_gin[1] = gins[1];
    }
    return launch_regional_any(d->dtype, g.luma3, prm, st);
}

int tfcfft_regional_spectra(const tfcfft_desc* d, const void* x, float* amp, float* pha, int fftshift, void* workspace,
                            size_t workspace_bytes, void* stream) {
    float* outs[2] = {amp, pha};
    return regional_spectra_common(d, x, nullptr, workspace, workspace_bytes, 1, fftshift, outs, nullptr, (cudaStream_t)stream);
}

int tfcfft_regional_spectra_bwd(const tfcfft_desc* d, const void* x, const float* grad_amp, const float* grad_pha, void* grad_x,
                                int fftshift, void* workspace, size_t workspace_bytes, void* stream) {
    if (!grad_x) return TFCFFT_ERR_NULL;
    const float* gins[2] = {grad_amp, grad_pha};
    return regional_spectra_common(d, x, grad_x, workspace, workspace_bytes, 2, fftshift, nullptr, gins, (cudaStream_t)stream);
}

size_t tfcfft_triplet_workspace_bytes(void) { return kTripletWsBytes; }

int tfcfft_patch_triplet(const tfcfft_desc* d, const void* fake, const void* real, const int32_t* negatives, float margin,
                         float eps, float* out, void* grad_fake, void* workspace, size_t workspace_bytes, void* stream) {
    int rc = validate_triplet(d, negatives);
    if (rc) return rc;
    if (!fake || !real || !out) return TFCFFT_ERR_NULL;
    if ((rc = check_grad_args(d, grad_fake))) return rc;
    if ((rc = check_alignment(d, fake, real, grad_fake))) return rc;
    if (!workspace || workspace_bytes < kTripletWsBytes || ((uintptr_t)workspace & 255)) return TFCFFT_ERR_WORKSPACE;
    const TripletParams tp = make_triplet_params(d, fake, real, negatives, margin, eps, out, grad_fake, workspace);
    cudaStream_t st = (cudaStream_t)stream;
    return launch_triplet_any(d->dtype, tp, st);
}

int tfcfft_temperature_triplet(const tfcfft_desc* d, const void* fake, const void* positive, const void* negative,
                               const int64_t* neg_stride, const float* lut, float margin, float eps, float* out, void* grad_fake,
                               void* workspace, size_t workspace_bytes, void* stream) {
    int rc = validate_temperature(d, neg_stride);
    if (rc) return rc;
    if (!fake || !positive || !negative || !lut || !out) return TFCFFT_ERR_NULL;
    if ((rc = check_grad_args(d, grad_fake))) return rc;
    if ((rc = check_alignment(d, fake, (d->flags & TFCFFT_TEMPS_POSITIVE) ? fake : positive, grad_fake))) return rc;
    if ((uintptr_t)negative % (4 * elem_size(d->dtype))) return TFCFFT_ERR_ALIGNMENT;
    if ((d->flags & TFCFFT_TEMPS_POSITIVE) && ((uintptr_t)positive % 16)) return TFCFFT_ERR_ALIGNMENT;
    if (!workspace || workspace_bytes < kTripletWsBytes || ((uintptr_t)workspace & 255)) return TFCFFT_ERR_WORKSPACE;
    const TripletParams tp = make_temperature_params(d, fake, positive, negative, neg_stride, lut, margin, eps, out, grad_fake, workspace);
    cudaStream_t st = (cudaStream_t)stream;
    return launch_triplet_any(d->dtype, tp, st);
}

int tfcfft_vectorize_temps(const tfcfft_desc* d, const void* x, const float* lut, float* out, void* stream) {
    int64_t st4[4] = {0, 0, 0, 1};
    int rc = validate_temperature(d, st4);
    if (rc) return rc;
    if (!x || !lut || !out) return TFCFFT_ERR_NULL;
    if ((uintptr_t)x % (4 * elem_size(d->dtype)) || (uintptr_t)out % 16) return TFCFFT_ERR_ALIGNMENT;
    TempsParams tp{};
    tp.x = x;
    for (int i = 0; i < 4; ++i) tp.xs[i] = d->fake_stride[i];
    tp.n = (int)d->n;
    tp.h = (int)d->h;
    tp.out = out;
    for (int i = 0; i < 256; ++i) tp.lut[i] = lut[i];
    return launch_temps_any(d->dtype, tp, (cudaStream_t)stream);
}

int tfcfft_grad_scale(void* dst, const void* src, int32_t dtype, int64_t numel, const float* dev_scale, float host_scale,
                      void* stream) {
    if (!dst || !src) return TFCFFT_ERR_NULL;
    if (numel <= 0) return numel == 0 ? 0 : TFCFFT_ERR_SHAPE;
    if (((uintptr_t)dst & 15) || ((uintptr_t)src & 15)) return TFCFFT_ERR_ALIGNMENT;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t es = elem_size(dtype);
    if (es == 0 || dtype == TFCFFT_U8) return TFCFFT_ERR_DTYPE;
    return launch_grad_scale_any(dtype, dst, src, numel, dev_scale, host_scale, st);
}

int tfcfft_grad_rescale(void* grad, int32_t dtype, int64_t numel, const float* grad_output_dev, float* applied_dev, void* workspace,
                        size_t workspace_bytes, void* stream) {
    if (!grad || !grad_output_dev || !applied_dev) return TFCFFT_ERR_NULL;
    if (numel <= 0) return numel == 0 ? 0 : TFCFFT_ERR_SHAPE;
    if ((uintptr_t)grad & 15) return TFCFFT_ERR_ALIGNMENT;
    if (!workspace || workspace_bytes < kWsHeader || ((uintptr_t)workspace & 255)) return TFCFFT_ERR_WORKSPACE;
    if (elem_size(dtype) == 0 || dtype == TFCFFT_U8) return TFCFFT_ERR_DTYPE;
    // the ticket of the rescale launch lives at byte 128 of the workspace header (the loss launches use byte 0)
    unsigned* ticket = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(workspace) + 128);
    return launch_grad_rescale_any(dtype, grad, numel, grad_output_dev, applied_dev, ticket, (cudaStream_t)stream);
}

void tfcfft_debug_trace(void* device_buffer) { g_trace.store((long long*)device_buffer); }

int64_t tfcfft_launch_count(void) { return g_launches.load(); }
void tfcfft_launch_count_reset(void) { g_launches.store(0); }

}  // extern "C"
