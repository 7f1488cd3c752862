/* tfcfft.h -- C ABI of the B200-native TFC-GAN frequency-domain loss path.
 *
 * The reference (nudro/TFC-GAN, Python) has no FFI: the path is a set of module-level Python
 * functions duplicated per training script.  This header is the boundary
 * the build creates for them; every entry point names the reference code it replaces
 * (paths relative to the reference checkout):
 *
 *   tfcfft_loss        TFC-GAN-FFT/TFCGAN_multigpu_patchFFT_16P.py:271-375  (FFT_Components,
 *                      fft_components, calculate_ffts: 16-patch loss)
 *                      TFC-GAN-FFT/TFCGAN_multigpu_patchFFT.py:498-511       (4-patch, mean)
 *                      TFC-GAN-FFT/TFCGAN_multigpu_patchFFT_experiment.py:317-339 (4-patch, sum)
 *                      TFC-GAN-FFT/TFCGAN_multigpu_globalFFT.py:494-499      (global)
 *                      TFC-STN/TFCGAN_STN21_Original_NewModel3_B2A.py:461-467 (global_fourier_loss)
 *                      TFC-GAN-FFT/Devcom_MagMSE.py:91-118                   (mse_spec, per image:
 *                      flags LOG_MAGNITUDE|FULL_SPECTRUM|DIST_MSE|NO_PHASE, per_image != NULL)
 *                      -- plus the backward pass the reference lacks (its loss is detached, SURVEY.md
 *                      section 0 fact 2): d loss / d fake, written to grad_fake in the same launch.
 *   tfcfft_grad_scale  the multiplication by grad_output that autograd performs for
 *                      scaler.scale(loss_G).backward() (TFCGAN_multigpu_patchFFT_16P.py:607-610)
 *
 * Conventions: plain pointers and sizes only; every device buffer (inputs, outputs, workspace) is
 * owned by the caller; the library allocates nothing on the device, keeps no pointer after a call
 * returns, never synchronises the host, and enqueues all work on the caller's stream.  Return
 * code 0 = OK, negative = tfcfft_status argument error (nothing was launched), positive =
 * cudaError_t.  No exceptions cross this boundary and there is no CPU fallback.
 */
#ifndef TFCFFT_H_
#define TFCFFT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TFCFFT_VERSION 200 /* major*100 + minor */

/* element type of fake / real / grad_fake */
enum tfcfft_dtype { TFCFFT_F32 = 0, TFCFFT_F16 = 1, TFCFFT_BF16 = 2, TFCFFT_U8 = 3 };

/* flags (default 0 = the reference training loss: luma, amplitude + phase, L1, half spectrum,
 * mean over patches) */
#define TFCFFT_CHANNELS_RGB  (1u << 0) /* one spectrum per channel instead of ITU-R 601 luma      */
#define TFCFFT_NO_PHASE      (1u << 1) /* amplitude term only: loss = weight * amp                */
#define TFCFFT_DIST_MSE      (1u << 2) /* squared difference instead of absolute difference       */
#define TFCFFT_PATCH_SUM     (1u << 3) /* sum over patches (fft_loss) instead of mean             */
#define TFCFFT_LOG_MAGNITUDE (1u << 4) /* log |F| (Devcom_MagMSE.py:105-106)                      */
#define TFCFFT_FULL_SPECTRUM (1u << 5) /* mean over the full P x P plane (fft2) not P x (P/2+1)   */
#define TFCFFT_QUANTIZE_U8   (1u << 6) /* reference-as-shipped input path: uint8 wrap + integer
                                          luma (patchFFT_16P.py:300); forward only               */
#define TFCFFT_GRAD_ACCUMULATE (1u << 8) /* grad_fake += instead of grad_fake = (all loss entry points): lets several
                                            loss terms share one gradient buffer without an extra add pass    */
#define TFCFFT_TEMPS_POSITIVE (1u << 9) /* tfcfft_temperature_triplet only: `positive` is an fp32 tensor of
                                           temperatures (the loader's T_B), not an image                  */
#define TFCFFT_USE_HALFLINE  (1u << 27) /* testing / A-B: 64x64 tiles on the half-line engine (two threads per line) */
#define TFCFFT_USE_PAIR      (1u << 28) /* testing: 64x64 tiles through the packed pair kernel            */
#define TFCFFT_USE_LINE      (1u << 29) /* testing: 64x64 tiles through the thread-per-line kernel        */
#define TFCFFT_FORCE_GENERIC (1u << 30) /* testing: bypass the packed 64x64 fast path                */
#define TFCFFT_FORCE_SPLIT   (1u << 31) /* testing: route P = 64 / 128 through the split path     */

enum tfcfft_status {
    TFCFFT_OK = 0,
    TFCFFT_ERR_NULL = -1,        /* null descriptor / required pointer                            */
    TFCFFT_ERR_STRUCT = -2,      /* struct_size mismatch (ABI version skew)                       */
    TFCFFT_ERR_DTYPE = -3,
    TFCFFT_ERR_SHAPE = -4,       /* H != W, C not in {1,3}, H % grid, patch side not in {16..512}  */
    TFCFFT_ERR_STRIDE = -5,      /* innermost stride != 1 or outer strides not multiples of 4      */
    TFCFFT_ERR_ALIGNMENT = -6,   /* base pointer not aligned to 4 elements                         */
    TFCFFT_ERR_FLAGS = -7,       /* unsupported flag combination                                   */
    TFCFFT_ERR_WORKSPACE = -8,   /* workspace null, misaligned or smaller than workspace_bytes     */
    TFCFFT_ERR_NO_GRADIENT = -9, /* gradient requested for QUANTIZE_U8 or a uint8 input            */
    TFCFFT_ERR_EMPTY = -10       /* N == 0                                                         */
};

typedef struct tfcfft_desc {
    uint32_t struct_size; /* sizeof(tfcfft_desc) */
    int32_t dtype;        /* tfcfft_dtype */
    int32_t grid;         /* patches per side: 1 global, 2 = 4-patch, 4 = 16-patch */
    uint32_t flags;
    int64_t n, c, h, w;      /* NCHW */
    int64_t fake_stride[4];  /* element strides; views such as B[:, :, 0:64, 64:128] are fine */
    int64_t real_stride[4];
    int64_t grad_stride[4];  /* ignored when grad_fake == NULL */
    float weight;            /* multiplies the loss and the gradient (1/100 at patchFFT_16P.py:607) */
    float input_scale;       /* x' = input_scale * x before the transform (ignored with QUANTIZE_U8) */
    /* Gradient-only scale: grad_fake = (d loss / d fake) * grad_scale_host * (*grad_scale_dev).  This is the factor
     * autograd would apply afterwards for scaler.scale(loss_G).backward() (patchFFT_16P.py:607-610): folding it into
     * the producing launch saves the separate scaling pass (2 of 5 tensor passes) and keeps fp16 gradients in range.
     * grad_scale_host == 0 is read as 1; grad_scale_dev may be NULL (a device float otherwise, read at kernel time:
     * no host synchronisation).  Neither touches the loss value. */
    float grad_scale_host;
    uint32_t reserved;       /* must be 0 */
    const float* grad_scale_dev;
} tfcfft_desc;

int tfcfft_version(void);
const char* tfcfft_strerror(int rc);

/* Host-only argument check; tfcfft_loss performs the same check first. */
int tfcfft_validate(const tfcfft_desc* d);

/* Bytes of caller-owned device scratch tfcfft_loss needs for `d` (0 if `d` is invalid). */
size_t tfcfft_workspace_bytes(const tfcfft_desc* d);
/* Same for tfcfft_spectra / tfcfft_spectra_bwd (they run on a different kernel geometry at P >= 128). */
size_t tfcfft_spectra_workspace_bytes(const tfcfft_desc* d);

/* Zeroes the workspace header.  Call once after allocating a workspace (and after any call that
 * returned a CUDA error); calls leave the tickets / counters of the header zeroed for the next call (its
 * scratch areas are rewritten before they are read).  A workspace must not be shared by calls that may run
 * concurrently on different streams. */
int tfcfft_workspace_init(void* workspace, size_t workspace_bytes, void* stream);

/* Loss (and, when grad_fake != NULL, d loss / d fake) in one pass over the inputs.
 *   out        device float[8]: weight * 1/2 (amp + pha)  [weight * amp with NO_PHASE], amp, pha,
 *              non-finite flag, the gradient scale that was applied (grad_scale_host * *grad_scale_dev; this
 *              element is what tfcfft_grad_rescale takes as `applied_dev`), 3 reserved
 *   per_image  device float[2*N] or NULL: per-image (amp, pha) terms; their mean over N is amp/pha
 *   grad_fake  device buffer of d->dtype laid out by grad_stride, or NULL for forward only
 * The reduction order is fixed: the loss is bit-stable from run to run. */
int tfcfft_loss(const tfcfft_desc* d, const void* fake, const void* real, float* out, float* per_image,
                void* grad_fake, void* workspace, size_t workspace_bytes, void* stream);

/* tfcfft_loss for the reference's 4-patch call convention, where the real quadrants arrive from the data loader as
 * four SEPARATE tensors: fft_loss(fake_B, B1, B2, B3, B4) (TFCGAN_multigpu_patchFFT_experiment.py:317-339; quadrants
 * B1..B4 = top-left, top-right, bottom-left, bottom-right: datasets_temp.py:76-118; same convention inline at
 * TFCGAN_multigpu_patchFFT.py:498-511).  d->grid must be 2; real_quadrants[i] is a [N, C, H/2, W/2] tensor of d->dtype
 * whose strides are d->real_stride (all four alike).  No concatenation copy is made. */
int tfcfft_loss_quads(const tfcfft_desc* d, const void* fake, const void* const real_quadrants[4], float* out,
                      float* per_image, void* grad_fake, void* workspace, size_t workspace_bytes, void* stream);

/* Materialised spectra: the differentiable counterpart of the reference's fft_components
 * (TFCGAN_multigpu_patchFFT_16P.py:293-319; global variant TFCGAN_multigpu_globalFFT.py:266-284) and of
 * make_spectra / sample_spectra (patchFFT_16P.py:284-289, 378-388).  d->grid must be 1 (the tensor IS the
 * patch).  Outputs are float [N][C'][P][W], W = P/2+1 (or P with FULL_SPECTRUM), C' = 1 (luma) or 3 (rgb);
 * with LOG_MAGNITUDE the amplitude is log|F|.  fftshift != 0 stores them in np.fft.fftshift order over both
 * axes like the reference (:279).  y and any output pointer may be NULL. */
int tfcfft_spectra(const tfcfft_desc* d, const void* x, const void* y, float* amp_x, float* pha_x, float* amp_y,
                   float* pha_y, int fftshift, void* workspace, size_t workspace_bytes, void* stream);

/* Backward of tfcfft_spectra for `x`: grad_x = d L / d x given d L / d amp_x and d L / d pha_x
 * (either may be NULL); d->grad_stride describes grad_x. */
int tfcfft_spectra_bwd(const tfcfft_desc* d, const void* x, const float* grad_amp, const float* grad_pha,
                       void* grad_x, int fftshift, void* workspace, size_t workspace_bytes, void* stream);

/* Regional FFT loss on the reference's two 100 x 256 bands ("hair" rows 0..99, "eyes" rows 100..199 of a 256 x 256
 * image): regional_fft_loss, TFCGAN_multigpu_patchFFT_withregion_FFT.py:353-402 -- per band the grey -> rfft2 ->
 * abs / arctan2 -> nn.L1Loss pipeline of the patch losses on a 100 x 129 half spectrum, the two bands summed,
 * 1/2 (amp + pha).  Arguments as tfcfft_loss (d->grid is ignored, H = W = 256 required; flags: CHANNELS_RGB,
 * NO_PHASE, DIST_MSE, QUANTIZE_U8); fused forward + backward; out[1], out[2] are the summed amp / pha terms. */
size_t tfcfft_regional_workspace_bytes(const tfcfft_desc* d);
int tfcfft_regional_loss(const tfcfft_desc* d, const void* fake, const void* real, float* out, float* per_image,
                         void* grad_fake, void* workspace, size_t workspace_bytes, void* stream);

/* Materialised spectra of the two bands -- the reference's inner reg_fft (withregion_FFT.py:358-371) -- and their
 * backward, so that other criteria (e.g. the KLDivLoss variant, ..._withregion_FFT_KL.py:398-414) can be applied
 * through autograd.  amp / pha: device float [N][C'][2 bands][100][129]; fftshift != 0 stores them fftshift-ed over
 * both axes like the reference (:253).  Either output (or incoming gradient) may be NULL. */
int tfcfft_regional_spectra(const tfcfft_desc* d, const void* x, float* amp, float* pha, int fftshift, void* workspace,
                            size_t workspace_bytes, void* stream);
int tfcfft_regional_spectra_bwd(const tfcfft_desc* d, const void* x, const float* grad_amp, const float* grad_pha,
                                void* grad_x, int fftshift, void* workspace, size_t workspace_bytes, void* stream);

/* Patch triplet loss of the generator step, forward + backward in one streaming pass (the first "next" row of
 * the hot-path scope table).  Replaces, for all g*g patches at once,
 *     triplet_loss(fake_B_i, B_i, random_patches[k_i])      with nn.TripletMarginLoss(margin, p=2)
 * (TFCGAN_multigpu_patchFFT_16P.py:75, :558-583; 4-patch copies TFCGAN_multigpu_patchFFT.py:474-484 and
 * TFCGAN_multigpu_globalFFT.py:470-480): distance = || a - b + eps ||_2 over one patch row of one channel,
 * loss = mean over (n, c, patch, row) of max(margin + d(fake_i, real_i) - d(fake_i, real_{k_i}), 0), which equals
 * the reference's 1/g^2 * sum of per-patch means.
 *   d           shape / dtype / strides / weight as for tfcfft_loss; input_scale is ignored; C in 1..4;
 *               flags: 0 or TFCFFT_GRAD_ACCUMULATE
 *   negatives   HOST int32[grid*grid]: k_i, the row-major index of the real patch used as negative for patch i
 *   out         device float[4]: weight*loss, loss, fraction of rows with an active hinge, 0
 *   grad_fake   device buffer (d->dtype, d->grad_stride) receiving weight * d loss / d fake, or NULL
 *   workspace   tfcfft_triplet_workspace_bytes() bytes, 256-byte aligned, header zeroed once
 *               (tfcfft_workspace_init); one workspace per concurrently used stream */
size_t tfcfft_triplet_workspace_bytes(void);
int tfcfft_patch_triplet(const tfcfft_desc* d, const void* fake, const void* real, const int32_t* negatives,
                         float margin, float eps, float* out, void* grad_fake, void* workspace,
                         size_t workspace_bytes, void* stream);

/* Temperature triplet loss of the generator step (second "next" row of the scope table), one streaming pass:
 *     criterion_temp(vectorize_temps(fake_B), TB, vectorize_temps(B_tf))      nn.TripletMarginLoss(margin, p=2)
 * (TFCGAN_multigpu_patchFFT_16P.py:80, :254-268, :585-595; TempVector_PyTorch datasets_temp.py:14-35).  A pixel's
 * temperature is lut[uint8(red channel)], the triplet distance runs over one image row.
 *   d           shape / dtype / fake_stride / real_stride (= strides of `positive`) / weight; grid is ignored;
 *               flags: TFCFFT_QUANTIZE_U8 = the reference as shipped (uint8 wrap like ToPILImage + table gather,
 *               forward only); without it the differentiable variant lut[0] + (lut[255]-lut[0])/255 * input_scale*x;
 *               TFCFFT_TEMPS_POSITIVE: `positive` already holds temperatures (fp32, the loader's T_B);
 *               TFCFFT_GRAD_ACCUMULATE
 *   negative    the augmented real batch (B_tf), same dtype as fake, strides in neg_stride[4]
 *   lut         HOST float[256] (np.linspace(24, 38, 256) upstream)
 *   grad_fake   NULL or weight * d loss / d fake; ONLY channel 0 is written (the others receive no gradient:
 *               zero the buffer first or use TFCFFT_GRAD_ACCUMULATE)
 *   workspace   as for tfcfft_patch_triplet */
int tfcfft_temperature_triplet(const tfcfft_desc* d, const void* fake, const void* positive, const void* negative,
                               const int64_t* neg_stride, const float* lut, float margin, float eps, float* out,
                               void* grad_fake, void* workspace, size_t workspace_bytes, void* stream);

/* Temperature map alone: out[n][0][y][x] = lut[uint8(x[n][0][y][x])] (fp32, contiguous [N,1,H,W]) -- the reference's
 * vectorize_temps (patchFFT_16P.py:260-268).  d describes x through fake_stride; no workspace. */
int tfcfft_vectorize_temps(const tfcfft_desc* d, const void* x, const float* lut, float* out, void* stream);

/* dst[i] = src[i] * host_scale * (*dev_scale)   (dev_scale may be NULL; dst may equal src);
 * numel elements of `dtype`, both 16-byte aligned. */
int tfcfft_grad_scale(void* dst, const void* src, int32_t dtype, int64_t numel, const float* dev_scale,
                      float host_scale, void* stream);

/* In-place correction of a gradient that was produced with an expected scale folded in (grad_scale_host /
 * grad_scale_dev above) once autograd's actual grad_output is known: grad *= (*grad_output_dev) / (*applied_dev), after
 * which *applied_dev = *grad_output_dev (so a repeated backward pass stays correct).  When the two scalars are equal
 * -- the normal case -- the kernel exits after reading them: no pass over the tensor.  `applied_dev` is a caller-owned
 * device float that the caller initialised with grad_scale_host * (*grad_scale_dev). */
int tfcfft_grad_rescale(void* grad, int32_t dtype, int64_t numel, const float* grad_output_dev, float* applied_dev,
                        void* workspace, size_t workspace_bytes, void* stream);

/* DEBUG / profiling aid, not part of the stable surface: while `device_buffer` is non-NULL the 64x64 line kernel
 * and the sub-tile launches record the global nanosecond timer at their stage boundaries for the first 6 work
 * units of every CTA (16 int64 per unit; [15] = 1 when the unit was traced; the packed pair kernel records
 * clock64).  Read by tools/trace_line.py and tools/trace_sub.py.  Pass NULL to switch off. */
void tfcfft_debug_trace(void* device_buffer);

/* Number of kernels this library has launched in this process since the last reset. */
int64_t tfcfft_launch_count(void);
void tfcfft_launch_count_reset(void);

#ifdef __cplusplus
}
#endif
#endif /* TFCFFT_H_ */
