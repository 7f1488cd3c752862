// tfcfft_api.cu -- the C-ABI entry points of libtfcfft.so (include/tfcfft.h).
//
// Host side only validates, builds the kernel parameter block and enqueues launches on the
// caller's stream.  No device allocation, no host synchronisation, no fallback of any kind.
#include <atomic>
#include <cstdio>

#include "host_common.h"
#include "spectral_kernels.cuh"

using namespace tfcfft;

namespace {

std::atomic<long long> g_launches{0};
std::atomic<long long*> g_trace{nullptr};  // debug only (tfcfft_debug_trace)

struct DeviceInfo {
    int sms = 0;
};
DeviceInfo device_info() {
    // queried per call: cheap (cached by the runtime) and keeps the library free of global state
    DeviceInfo di;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&di.sms, cudaDevAttrMultiProcessorCount, dev);
    return di;
}

#define TFC_LAUNCH_CHECK()                       \
    do {                                         \
        cudaError_t e__ = cudaGetLastError();    \
        if (e__ != cudaSuccess) return (int)e__; \
    } while (0)

template <typename K>
int set_smem(K kernel, size_t bytes) {
    if (bytes > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) return (int)e;
    }
    return 0;
}

template <int P, typename T, bool LUMA3>
int launch_resident(const Params& prm, cudaStream_t st) {
    auto kernel = resident_kernel<P, T, LUMA3>;
    constexpr size_t smem = ResidentCfg<P>::SMEM;
    constexpr int nt = ResidentCfg<P>::NT;
    if (int rc = set_smem(kernel, smem)) return rc;
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, nt, smem);
    if (e != cudaSuccess) return (int)e;
    if (per_sm < 1) per_sm = 1;
    const long long cap = (long long)device_info().sms * per_sm;  // persistent: one wave
    const int grid = (int)(prm.tiles_total < cap ? prm.tiles_total : cap);
    kernel<<<grid, nt, smem, st>>>(prm);
    g_launches++;
    TFC_LAUNCH_CHECK();
    return 0;
}

template <int P, typename T, bool LUMA3>
int launch_pair(const Params& prm, cudaStream_t st) {
    auto kernel = pair_kernel<P, T, LUMA3>;
    constexpr size_t smem = PairCfg<P>::SMEM;
    constexpr int nt = PairCfg<P>::NT;
    if (int rc = set_smem(kernel, smem)) return rc;
    const long long npairs = ((long long)prm.tiles_total + 1) / 2;
    const long long cap = device_info().sms;  // persistent, warp-specialised: one CTA per SM
    const int grid = (int)(npairs < cap ? npairs : cap);
    kernel<<<grid, nt, smem, st>>>(prm);
    g_launches++;
    TFC_LAUNCH_CHECK();
    return 0;
}

// launch with the programmatic-stream-serialization attribute (spectral_kernels.cuh: pdl_wait / pdl_release)
template <class K, class P>
cudaError_t launch_pdl(K kernel, int grid, int block, size_t smem, cudaStream_t st, const P& prm) {
    static const bool off = getenv("TFCFFT_NO_PDL") != nullptr;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = off ? 0 : 1;
    return cudaLaunchKernelEx(&cfg, kernel, prm);
}

template <typename T, bool LUMA3>
int launch_line(const Params& prm, cudaStream_t st) {
    auto kernel = line_kernel<T, LUMA3>;
    constexpr size_t smem = LineCfg::SMEM;
    if (int rc = set_smem(kernel, smem)) return rc;
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, LineCfg::NT, smem);
    if (e != cudaSuccess) return (int)e;
    if (per_sm < 1) per_sm = 1;
    const long long cap = (long long)device_info().sms * per_sm;
    const int grid = (int)(prm.tiles_total < cap ? prm.tiles_total : cap);
    if (cudaError_t e2 = launch_pdl(kernel, grid, LineCfg::NT, smem, st, prm)) return (int)e2;
    g_launches++;
    return 0;
}

template <typename T, bool LUMA3>
int launch_sub(Params prm, cudaStream_t st) {
    auto kf = sub_fwd_kernel<T, LUMA3>;
    auto ki = sub_inv_kernel<T, LUMA3>;
    if (int rc = set_smem(kf, SubCfg::SMEM_FWD)) return rc;
    if (int rc = set_smem(ki, SubCfg::SMEM_INV)) return rc;
    static int per_sm_f = 0, per_sm_i = 0;
    if (!per_sm_f) {
        int f = 0, i = 0;
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&f, kf, SubCfg::NT_FWD, SubCfg::SMEM_FWD);
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&i, ki, SubCfg::NT_INV, SubCfg::SMEM_INV);
        if (e != cudaSuccess) return (int)e;
        per_sm_i = i < 1 ? 1 : i;
        per_sm_f = f < 1 ? 1 : f;
    }
    const int D = prm.sub_d, npp = D * D / 2;
    const int sms = device_info().sms;
    for (int base = 0; base < prm.tiles_total; base += prm.chunk_tiles) {
        prm.tile_base = base;
        prm.chunk_now = prm.tiles_total - base < prm.chunk_tiles ? prm.tiles_total - base : prm.chunk_tiles;
        const int units = prm.chunk_now * npp;
        const int grid_f = units < sms * per_sm_f ? units : sms * per_sm_f;
        const int grid_i = units < sms * per_sm_i ? units : sms * per_sm_i;
        cudaError_t e;
        static const bool no_cluster = getenv("TFCFFT_NO_CLUSTER") != nullptr;
        if (D == 4 && !no_cluster) {  // 2-CTA clusters: full-sector loads, halves exchanged through DSMEM
            auto kf4 = sub_fwd4_kernel<T, LUMA3>;
            if (int rc = set_smem(kf4, SubCfg::SMEM_FWD)) return rc;
            int g4 = units < sms * per_sm_f ? units : sms * per_sm_f;
            g4 &= ~1;
            e = launch_pdl(kf4, g4, SubCfg::NT_FWD, SubCfg::SMEM_FWD, st, prm);
        } else {
            e = launch_pdl(kf, grid_f, SubCfg::NT_FWD, SubCfg::SMEM_FWD, st, prm);
        }
        if (e != cudaSuccess) return (int)e;
        g_launches++;
        e = D == 2 ? launch_pdl(combine_kernel<2>, prm.chunk_now * kCombineParts, kCombineThreads, 0, st, prm)
                   : launch_pdl(combine_kernel<4>, prm.chunk_now * kCombineParts, kCombineThreads, 0, st, prm);
        if (e != cudaSuccess) return (int)e;
        g_launches++;
        if (prm.grad) {
            if (D == 4 && !no_cluster) {
                auto ki4 = sub_inv4_kernel<T, LUMA3>;
                if (int rc = set_smem(ki4, SubCfg::SMEM_INV)) return rc;
                int g4 = units < sms * per_sm_i ? units : sms * per_sm_i;
                g4 &= ~1;
                e = launch_pdl(ki4, g4, SubCfg::NT_INV, SubCfg::SMEM_INV, st, prm);
            } else {
                e = launch_pdl(ki, grid_i, SubCfg::NT_INV, SubCfg::SMEM_INV, st, prm);
            }
            if (e != cudaSuccess) return (int)e;
            g_launches++;
        }
    }
    return 0;
}

template <int P, typename T, bool LUMA3>
int launch_split(Params prm, cudaStream_t st) {
    if constexpr (P >= 64) {
        using Sp = Split<P>;
        auto k1 = split_rows_fwd_kernel<P, T, LUMA3>;
        auto k2 = split_cols_kernel<P>;
        auto k3 = split_rows_inv_kernel<P, T, LUMA3>;
        if (int rc = set_smem(k1, SplitCfg<P>::SMEM_ROWS)) return rc;
        if (int rc = set_smem(k2, SplitCfg<P>::SMEM_COLS)) return rc;
        if (int rc = set_smem(k3, SplitCfg<P>::SMEM_ROWS)) return rc;
        for (int base = 0; base < prm.tiles_total; base += prm.chunk_tiles) {
            prm.tile_base = base;
            const int nt = prm.tiles_total - base < prm.chunk_tiles ? prm.tiles_total - base : prm.chunk_tiles;
            if (cudaError_t e = launch_pdl(k1, nt * Sp::ROW_SLABS, SplitCfg<P>::NT, SplitCfg<P>::SMEM_ROWS, st, prm)) return (int)e;
            g_launches++;
            if (cudaError_t e = launch_pdl(k2, nt * Sp::PARTS, SplitCfg<P>::NT, SplitCfg<P>::SMEM_COLS, st, prm)) return (int)e;
            g_launches++;
            if (prm.grad) {
                if (cudaError_t e = launch_pdl(k3, nt * Sp::ROW_SLABS, SplitCfg<P>::NT, SplitCfg<P>::SMEM_ROWS, st, prm)) return (int)e;
                g_launches++;
            }
        }
        return 0;
    } else {
        return TFCFFT_ERR_SHAPE;
    }
}

template <int P, typename T, bool LUMA3>
int launch(const Params& prm, bool split, cudaStream_t st) {
    if constexpr (P == 128 || P == 256) {
        if (prm.sub_d > 1) return launch_sub<T, LUMA3>(prm, st);
    }
    if (split) return launch_split<P, T, LUMA3>(prm, st);
    if constexpr (P <= 128) {
        if constexpr (P == 64) {
            if (pair_supported(prm)) {
                // measured (profiles/): the thread-per-line kernel wins on both luma (1.45 M vs 1.39 M img/s) and
                // single-channel tiles (0.80 M vs 0.60 M); the packed pair kernel stays selectable for A/B runs
                const bool line = !(prm.flags & TFCFFT_USE_PAIR);
                if (line) return launch_line<T, LUMA3>(prm, st);
                return launch_pair<P, T, LUMA3>(prm, st);
            }
        }
        return launch_resident<P, T, LUMA3>(prm, st);
    } else {
        return TFCFFT_ERR_SHAPE;
    }
}

template <int P, typename T>
int launch_l(const Params& prm, bool split, bool luma3, cudaStream_t st) {
    return luma3 ? launch<P, T, true>(prm, split, st) : launch<P, T, false>(prm, split, st);
}

template <int P>
int launch_t(const Params& prm, bool split, bool luma3, int dtype, cudaStream_t st) {
    switch (dtype) {
        case TFCFFT_F32: return launch_l<P, float>(prm, split, luma3, st);
        case TFCFFT_F16: return launch_l<P, __half>(prm, split, luma3, st);
        case TFCFFT_BF16: return launch_l<P, __nv_bfloat16>(prm, split, luma3, st);
        case TFCFFT_U8: return launch_l<P, uint8_t>(prm, split, luma3, st);
    }
    return TFCFFT_ERR_DTYPE;
}

template <typename T>
int launch_triplet(const TripletParams& tp, cudaStream_t st) {
    // float4 per lane: 8 lanes per patch row up to 128-pixel rows (the per-row scalar work -- row decoding, shuffles,
    // rsqrt -- is repeated by every lane of the group; ncu: 80 % issue-slot utilisation with 16 lanes x 1 float4)
    const int K = tp.p >= 128 ? 4 : tp.p >= 64 ? 2 : 1;
    const int gpw = 32 / (tp.p / (4 * K));
    const long long warps_needed = (tp.rows + gpw - 1) / gpw;
    long long blocks = (warps_needed + kTripletThreads / 32 - 1) / (kTripletThreads / 32);
    void (*kernel)(TripletParams) = K == 1 ? triplet_kernel<T, 1> : K == 2 ? triplet_kernel<T, 2> : triplet_kernel<T, 4>;
    if (K != 1 && K != 2 && K != 4) return TFCFFT_ERR_SHAPE;
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kTripletThreads, 0);
    if (e != cudaSuccess) return (int)e;
    // grid-stride warps over ~4 resident waves: measured 1.33 M img/s with exactly one wave, 1.45 M with three or more
    // (the block scheduler evens out the skew between warps that a single persistent wave keeps to the end)
    const long long cap = (long long)device_info().sms * (per_sm < 1 ? 1 : per_sm) * 4;
    if (blocks > cap) blocks = cap;
    if (blocks > kTripletMaxBlocks) blocks = kTripletMaxBlocks;
    if (cudaError_t e2 = launch_pdl(kernel, (int)blocks, kTripletThreads, 0, st, tp)) return (int)e2;
    g_launches++;
    return 0;
}

template <typename T, bool LUMA3>
int launch_regional(const Params& prm, cudaStream_t st) {
    auto kernel = regional_kernel<T, LUMA3>;
    if (int rc = set_smem(kernel, RegCfg::SMEM)) return rc;
    const int sms = device_info().sms;
    const int grid = prm.tiles_total < sms ? prm.tiles_total : sms;
    if (cudaError_t e = launch_pdl(kernel, grid, RegCfg::NT, RegCfg::SMEM, st, prm)) return (int)e;
    g_launches++;
    return 0;
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

int tfcfft_workspace_init(void* workspace, size_t workspace_bytes, void* stream) {
    if (!workspace || workspace_bytes < kWsHeader || ((uintptr_t)workspace & 255)) return TFCFFT_ERR_WORKSPACE;
    cudaError_t e = cudaMemsetAsync(workspace, 0, kWsHeader, (cudaStream_t)stream);
    return e == cudaSuccess ? 0 : (int)e;
}

static int dispatch(const Params& prm, const Geometry& g, int dtype, cudaStream_t st) {
    switch (g.p) {
        case 16: return launch_t<16>(prm, g.split, g.luma3, dtype, st);
        case 32: return launch_t<32>(prm, g.split, g.luma3, dtype, st);
        case 64: return launch_t<64>(prm, g.split, g.luma3, dtype, st);
        case 128: return launch_t<128>(prm, g.split, g.luma3, dtype, st);
        case 256: return launch_t<256>(prm, g.split, g.luma3, dtype, st);
        case 512: return launch_t<512>(prm, g.split, g.luma3, dtype, st);
    }
    return TFCFFT_ERR_SHAPE;
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

int tfcfft_loss(const tfcfft_desc* d, const void* fake, const void* real, float* out, float* per_image, void* grad_fake,
                void* workspace, size_t workspace_bytes, void* stream) {
    Geometry g;
    int rc = validate_desc(d, &g);
    if (rc) return rc;
    if (!fake || !real || !out) return TFCFFT_ERR_NULL;
    if ((rc = check_grad_args(d, grad_fake))) return rc;
    if ((rc = check_alignment(d, fake, real, grad_fake))) return rc;
    if (!workspace || workspace_bytes < g.ws_bytes || ((uintptr_t)workspace & 255)) return TFCFFT_ERR_WORKSPACE;
    Params prm = make_params(d, g, fake, real, grad_fake, out, per_image, workspace);
    prm.trace = g_trace.load();
    cudaStream_t st = (cudaStream_t)stream;
    return dispatch(prm, g, d->dtype, st);
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
    switch (d->dtype) {
        case TFCFFT_F32: return g.luma3 ? launch_regional<float, true>(prm, st) : launch_regional<float, false>(prm, st);
        case TFCFFT_F16: return g.luma3 ? launch_regional<__half, true>(prm, st) : launch_regional<__half, false>(prm, st);
        case TFCFFT_BF16: return g.luma3 ? launch_regional<__nv_bfloat16, true>(prm, st) : launch_regional<__nv_bfloat16, false>(prm, st);
        case TFCFFT_U8: return g.luma3 ? launch_regional<uint8_t, true>(prm, st) : launch_regional<uint8_t, false>(prm, st);
    }
    return TFCFFT_ERR_DTYPE;
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
    switch (d->dtype) {
        case TFCFFT_F32: return g.luma3 ? launch_regional<float, true>(prm, st) : launch_regional<float, false>(prm, st);
        case TFCFFT_F16: return g.luma3 ? launch_regional<__half, true>(prm, st) : launch_regional<__half, false>(prm, st);
        case TFCFFT_BF16: return g.luma3 ? launch_regional<__nv_bfloat16, true>(prm, st) : launch_regional<__nv_bfloat16, false>(prm, st);
        case TFCFFT_U8: return g.luma3 ? launch_regional<uint8_t, true>(prm, st) : launch_regional<uint8_t, false>(prm, st);
    }
    return TFCFFT_ERR_DTYPE;
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
    switch (d->dtype) {
        case TFCFFT_F32: return launch_triplet<float>(tp, st);
        case TFCFFT_F16: return launch_triplet<__half>(tp, st);
        case TFCFFT_BF16: return launch_triplet<__nv_bfloat16>(tp, st);
        case TFCFFT_U8: return launch_triplet<uint8_t>(tp, st);
    }
    return TFCFFT_ERR_DTYPE;
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
    switch (d->dtype) {
        case TFCFFT_F32: return launch_triplet<float>(tp, st);
        case TFCFFT_F16: return launch_triplet<__half>(tp, st);
        case TFCFFT_BF16: return launch_triplet<__nv_bfloat16>(tp, st);
        case TFCFFT_U8: return launch_triplet<uint8_t>(tp, st);
    }
    return TFCFFT_ERR_DTYPE;
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
    const long long total4 = (long long)d->n * d->h * (d->h / 4);
    long long blocks = (total4 + 255) / 256;
    const long long cap = (long long)device_info().sms * 16;
    if (blocks > cap) blocks = cap;
    cudaStream_t st = (cudaStream_t)stream;
    switch (d->dtype) {
        case TFCFFT_F32: temps_kernel<float><<<(int)blocks, 256, 0, st>>>(tp); break;
        case TFCFFT_F16: temps_kernel<__half><<<(int)blocks, 256, 0, st>>>(tp); break;
        case TFCFFT_BF16: temps_kernel<__nv_bfloat16><<<(int)blocks, 256, 0, st>>>(tp); break;
        case TFCFFT_U8: temps_kernel<uint8_t><<<(int)blocks, 256, 0, st>>>(tp); break;
        default: return TFCFFT_ERR_DTYPE;
    }
    g_launches++;
    TFC_LAUNCH_CHECK();
    return 0;
}

int tfcfft_grad_scale(void* dst, const void* src, int32_t dtype, int64_t numel, const float* dev_scale, float host_scale,
                      void* stream) {
    if (!dst || !src) return TFCFFT_ERR_NULL;
    if (numel <= 0) return numel == 0 ? 0 : TFCFFT_ERR_SHAPE;
    if (((uintptr_t)dst & 15) || ((uintptr_t)src & 15)) return TFCFFT_ERR_ALIGNMENT;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t es = elem_size(dtype);
    if (es == 0 || dtype == TFCFFT_U8) return TFCFFT_ERR_DTYPE;
    const long long nvec = numel / (16 / (long long)es) + 1;
    long long blocks = (nvec + 255) / 256;
    const long long cap = (long long)device_info().sms * 16;
    if (blocks > cap) blocks = cap;
    switch (dtype) {
        case TFCFFT_F32:
            grad_scale_kernel<float><<<(int)blocks, 256, 0, st>>>((float*)dst, (const float*)src, numel, dev_scale, host_scale);
            break;
        case TFCFFT_F16:
            grad_scale_kernel<__half><<<(int)blocks, 256, 0, st>>>((__half*)dst, (const __half*)src, numel, dev_scale, host_scale);
            break;
        case TFCFFT_BF16:
            grad_scale_kernel<__nv_bfloat16><<<(int)blocks, 256, 0, st>>>((__nv_bfloat16*)dst, (const __nv_bfloat16*)src, numel, dev_scale, host_scale);
            break;
    }
    g_launches++;
    TFC_LAUNCH_CHECK();
    return 0;
}

void tfcfft_debug_trace(void* device_buffer) { g_trace.store((long long*)device_buffer); }

int64_t tfcfft_launch_count(void) { return g_launches.load(); }
void tfcfft_launch_count_reset(void) { g_launches.store(0); }

}  // extern "C"
