// launchers.h -- host entry points of the kernel translation units (k_*.cu), called by tfcfft_api.cu.
// Every function validates nothing: arguments were checked by the C-ABI layer.  Return 0, a negative
// TFCFFT_ERR_* or a positive cudaError_t.
#pragma once
#include <mutex>
#include "host_common.h"
#include "kernel_common.cuh"

namespace tfcfft {

// temperature map alone (vectorize_temps): red channel -> uint8 like ToPILImage -> table
struct TempsParams {
    const void* x;
    long long xs[4];
    int n, h;
    float* out;
    float lut[256];
};

// The four FFT kernel families are compiled once per element type (k_<family>.cu with -DTFC_DT=0..3: the build
// parallelises over 16 objects instead of 4); each object exports the entry points of its dtype.
#define TFC_DT_LIST(X) X(0, float, f32) X(1, __half, f16) X(2, __nv_bfloat16, bf16) X(3, uint8_t, u8)
#define TFC_DECL_DT(n, T, sfx)                                                                                  \
    int launch_resident_##sfx(int p, bool luma3, const Params& prm, cudaStream_t st); /* k_resident.cu, P <= 128 */ \
    int launch_pair_##sfx(bool luma3, const Params& prm, cudaStream_t st);            /* k_resident.cu, P == 64  */ \
    int launch_line_##sfx(bool luma3, const Params& prm, cudaStream_t st);            /* k_line.cu, P == 64      */ \
    int launch_sub_##sfx(bool luma3, const Params& prm, cudaStream_t st);             /* k_sub.cu, P = 128 / 256 */ \
    int launch_split_##sfx(int p, bool luma3, const Params& prm, cudaStream_t st);    /* k_split.cu, P >= 64     */
TFC_DT_LIST(TFC_DECL_DT)
#undef TFC_DECL_DT

#define TFC_ANY(family, ARGS_DECL, ARGS)                              \
    inline int launch_##family##_any ARGS_DECL {                      \
        switch (dtype) {                                              \
            case TFCFFT_F32: return launch_##family##_f32 ARGS;       \
            case TFCFFT_F16: return launch_##family##_f16 ARGS;       \
            case TFCFFT_BF16: return launch_##family##_bf16 ARGS;     \
            case TFCFFT_U8: return launch_##family##_u8 ARGS;         \
        }                                                             \
        return TFCFFT_ERR_DTYPE;                                      \
    }
TFC_ANY(resident, (int p, int dtype, bool luma3, const Params& prm, cudaStream_t st), (p, luma3, prm, st))
TFC_ANY(pair, (int dtype, bool luma3, const Params& prm, cudaStream_t st), (luma3, prm, st))
TFC_ANY(line, (int dtype, bool luma3, const Params& prm, cudaStream_t st), (luma3, prm, st))
TFC_ANY(sub, (int dtype, bool luma3, const Params& prm, cudaStream_t st), (luma3, prm, st))
TFC_ANY(split, (int p, int dtype, bool luma3, const Params& prm, cudaStream_t st), (p, luma3, prm, st))
#undef TFC_ANY

// dtype-independent launches of the multi-launch pipelines (defined in the TFC_DT == 0 object of their unit)
cudaError_t launch_combine(int d, int grid, const Params& prm, cudaStream_t st);            // k_sub.cu
struct Lanes {  // auxiliary stream + fork / join events of the two-lane chunk schedule (k_sub.cu), one per device
    cudaStream_t aux;
    cudaEvent_t fork, join;
    std::mutex mu;
};
Lanes* lanes_get();                                                                          // k_sub.cu
int split_cols_facts(int p);                                                                 // k_split.cu
cudaError_t launch_split_cols(int p, int grid, const Params& prm, cudaStream_t st);          // k_split.cu

int launch_triplet_any(int dtype, const TripletParams& tp, cudaStream_t st);                // k_misc.cu
int launch_regional_any(int dtype, bool luma3, const Params& prm, cudaStream_t st);         // k_misc.cu
int launch_temps_any(int dtype, const TempsParams& tp, cudaStream_t st);                    // k_misc.cu
int launch_grad_scale_any(int dtype, void* dst, const void* src, long long numel, const float* dev_scale, float host_scale,
                          cudaStream_t st);                                                 // k_misc.cu
int launch_grad_rescale_any(int dtype, void* grad, long long numel, const float* go_dev, float* applied_dev, unsigned* ticket,
                            cudaStream_t st);                                               // k_misc.cu

// element type of the object being compiled
#ifdef TFC_DT
#if TFC_DT == 0
#define TFC_T float
#define TFC_FN(name) name##_f32
#elif TFC_DT == 1
#define TFC_T __half
#define TFC_FN(name) name##_f16
#elif TFC_DT == 2
#define TFC_T __nv_bfloat16
#define TFC_FN(name) name##_bf16
#elif TFC_DT == 3
#define TFC_T uint8_t
#define TFC_FN(name) name##_u8
#else
#error "TFC_DT must be 0..3"
#endif
#endif

// dtype x luma dispatch used by k_misc.cu
#define TFC_DISPATCH_T_L(dtype, luma3, CALL)                                                        \
    switch (dtype) {                                                                                \
        case TFCFFT_F32: return (luma3) ? CALL(float, true) : CALL(float, false);                   \
        case TFCFFT_F16: return (luma3) ? CALL(__half, true) : CALL(__half, false);                 \
        case TFCFFT_BF16: return (luma3) ? CALL(__nv_bfloat16, true) : CALL(__nv_bfloat16, false);  \
        case TFCFFT_U8: return (luma3) ? CALL(uint8_t, true) : CALL(uint8_t, false);                \
    }                                                                                               \
    return TFCFFT_ERR_DTYPE

}  // namespace tfcfft
