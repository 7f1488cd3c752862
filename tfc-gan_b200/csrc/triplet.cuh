// triplet.cuh -- patch triplet loss of the generator step, fused forward + backward.
//
// Reference: TFC-GAN-FFT/TFCGAN_multigpu_patchFFT_16P.py:75 (`nn.TripletMarginLoss(margin=1.0, p=2)`) and :558-583
// (anchor = fake patch i, positive = real patch i, negative = a randomly drawn real patch k_i, mean over the 16
// patches); 4-patch copies in TFCGAN_multigpu_patchFFT.py:474-480 and TFCGAN_multigpu_globalFFT.py:470-476.
// torch semantics: d(x1, x2) = || x1 - x2 + eps ||_2 over the LAST dimension (one patch row of one channel),
// loss = mean over (n, c, patch, row) of max(margin + d(a, p) - d(a, n), 0).
//
// One streaming pass: a group of lanes owns one patch row, reads anchor / positive / negative once (16-byte
// loads, streaming hints), reduces the two squared distances with warp shuffles, and -- still holding the
// differences in registers -- writes d loss / d fake for that row.  4 tensor passes (3 reads + 1 write; +1 read
// with TFCFFT_GRAD_ACCUMULATE) against ~20 for the eager op chain; the loss is reduced deterministically
// (per-CTA partials, last CTA sums them in a fixed order in double).
#pragma once
#include "spectral_core.cuh"

namespace tfcfft {

struct TripletParams {
    const void *fake, *real;
    void* grad;
    long long fs[4], rs[4], gs[4];
    int n, c, h, grid, p;
    int lg_g, lg_h, lg_p;  // grid, h and p are powers of two: row decoding is shifts and masks
    int neg[16];       // negative patch index for each patch (row-major patch order)
    // temperature variant (tfcfft_temperature_triplet): the negative is a third tensor, values pass through the
    // temperature map first, only channel 0 takes part (c == 1)
    const void* neg_src;  // NULL: negative = tile neg[i] of `real`
    long long ns[4];
    int mode;          // 0 raw values; 1 lut[uint8(x)] (reference as shipped, no gradient); 2 lin_a + lin_b * x
    int pos_f32;       // `real` is an fp32 tensor that already holds temperatures (the loader's T_B)
    float lin_a, lin_b;
    float lut[256];
    float margin, eps;
    float weight;      // out[0] = weight * mean hinge
    float coef;        // weight / rows: gradient scale of one active row
    int accumulate;    // grad += instead of grad =
    long long rows;    // N * C * H * grid patch rows
    float* partials;   // [blocks][2]: (sum of hinge, number of active rows)
    unsigned* counter;
    float* out;        // [4]: weight * loss, loss, active fraction, 0
};

constexpr int kTripletThreads = 256;
constexpr int kTripletRowsInFlight = 1;  // rows per lane group per pass (device kernel)

// product rounded on its own (never contracted into an FMA): a patch that draws itself as its negative must get a
// gradient of exactly zero, like the reference, and `ca * dp - cn * dn` only cancels exactly with two rounded products
TFC_HD float mul_rn(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fmul_rn(a, b);
#else
    volatile float r = a * b;
    return r;
#endif
}

TFC_HD float inv_sqrt(float x) {
#ifdef __CUDA_ARCH__
    return rsqrtf(x);
#else
    return 1.0f / sqrtf(x);
#endif
}

struct SerialReduce {
    TFC_HD float operator()(float v) const { return v; }
};

// temperature of one pixel value (TFCGAN_multigpu_patchFFT_16P.py:257-268, datasets_temp.py:14-35): the reference
// quantises to uint8 like ToPILImage and gathers from a 256-entry table; the differentiable variant is the
// table's linear law on the unquantised value
template <typename T>
TFC_HD float temp_value(const TripletParams& tp, float raw) {
    if (tp.mode == 1) return tp.lut[IO<T>::quant(raw)];
    return fmaf(raw, tp.lin_b, tp.lin_a);
}

// One patch row by a group of `lpr` lanes, K float4 per lane (lane `l` takes float4 l, l + lpr, ...), in two steps so
// that a warp can have the loads of several rows in flight before it reduces the first.
template <int K>
struct TripletRow {
    float dp[K][4], dn[K][4];  // anchor - positive + eps, anchor - negative + eps
    long long goff;            // element offset of the row segment in grad
};

template <typename T, int K>
TFC_HD void triplet_row_load(const TripletParams& tp, long long row, int l, int lpr, TripletRow<K>& tr) {
    const int g = tp.grid, P = tp.p;
    const int px = (int)(row & (g - 1));
    const int y = (int)((row >> tp.lg_g) & (tp.h - 1));
    const unsigned nc = (unsigned)(row >> (tp.lg_g + tp.lg_h));  // n * C + c < 2^27
    const int n = (int)(nc / (unsigned)tp.c), c = (int)(nc - (unsigned)n * (unsigned)tp.c);
    const int py = y >> tp.lg_p, yin = y & (P - 1);
    const int k = tp.neg[(py << tp.lg_g) + px], ky = k >> tp.lg_g, kx = k & (g - 1);
    const T* fp = static_cast<const T*>(tp.fake) + n * tp.fs[0] + c * tp.fs[1] + (long long)y * tp.fs[2] + px * P;
    const T* pp = static_cast<const T*>(tp.real) + n * tp.rs[0] + c * tp.rs[1] + (long long)y * tp.rs[2] + px * P;
    const T* qp = tp.neg_src == nullptr
                      ? static_cast<const T*>(tp.real) + n * tp.rs[0] + c * tp.rs[1] + (long long)(ky * P + yin) * tp.rs[2] + kx * P
                      : static_cast<const T*>(tp.neg_src) + n * tp.ns[0] + c * tp.ns[1] + (long long)y * tp.ns[2] + px * P;
    tr.goff = n * tp.gs[0] + c * tp.gs[1] + (long long)y * tp.gs[2] + px * P;
    float f[K][4];
#pragma unroll
    for (int j = 0; j < K; ++j) {
        const int x = 4 * (l + j * lpr);
        IO<T>::load4(fp + x, f[j]);
        if (tp.pos_f32) IO<float>::load4(reinterpret_cast<const float*>(tp.real) + (pp - static_cast<const T*>(tp.real)) + x, tr.dp[j]);
        else IO<T>::load4(pp + x, tr.dp[j]);
        IO<T>::load4(qp + x, tr.dn[j]);
    }
    if (tp.mode != 0) {
#pragma unroll
        for (int j = 0; j < K; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                f[j][i] = temp_value<T>(tp, f[j][i]);
                if (!tp.pos_f32) tr.dp[j][i] = temp_value<T>(tp, tr.dp[j][i]);
                tr.dn[j][i] = temp_value<T>(tp, tr.dn[j][i]);
            }
    }
#pragma unroll
    for (int j = 0; j < K; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            tr.dp[j][i] = (f[j][i] - tr.dp[j][i]) + tp.eps;
            tr.dn[j][i] = (f[j][i] - tr.dn[j][i]) + tp.eps;
        }
}

template <typename T, int K, class Reduce>
TFC_HD void triplet_row_finish(const TripletParams& tp, const TripletRow<K>& tr, int l, int lpr, const Reduce& reduce,
                               float& loss_acc, float& act_acc) {
    float sap = 0.f, san = 0.f;
#pragma unroll
    for (int j = 0; j < K; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            sap = fmaf(tr.dp[j][i], tr.dp[j][i], sap);
            san = fmaf(tr.dn[j][i], tr.dn[j][i], san);
        }
    sap = reduce(sap);
    san = reduce(san);
    // one reciprocal square root per distance (MUFU) instead of sqrt + divide: every lane of the group repeats this
    // scalar tail, and IEEE sqrt / divide sequences made it a third of the kernel's instructions
    const float iap = sap > 0.f ? inv_sqrt(sap) : 0.f, ian = san > 0.f ? inv_sqrt(san) : 0.f;
    const float dap = sap * iap, dan = san * ian;
    const float hinge = tp.margin + dap - dan;
    const bool active = hinge >= 0.f;  // torch clamp_min backward passes the gradient where input >= min
    if (l == 0) {
        loss_acc += active ? hinge : 0.f;
        act_acc += active ? 1.f : 0.f;
    }
    if (tp.grad != nullptr) {
        const float ca = active ? tp.coef * iap : 0.f;
        const float cn = active ? tp.coef * ian : 0.f;
        T* gp = static_cast<T*>(tp.grad) + tr.goff;
#pragma unroll
        for (int j = 0; j < K; ++j) {
            const int x = 4 * (l + j * lpr);
            float v[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) v[i] = mul_rn(ca, tr.dp[j][i]) - mul_rn(cn, tr.dn[j][i]);
            if (tp.accumulate) {
                float old[4];
                IO<T>::load4(gp + x, old);
#pragma unroll
                for (int i = 0; i < 4; ++i) v[i] += old[i];
            }
            IO<T>::store4(gp + x, v);
        }
    }
}

template <typename T, int K, class Reduce>
TFC_HD void triplet_row(const TripletParams& tp, long long row, int l, int lpr, const Reduce& reduce, float& loss_acc,
                        float& act_acc) {
    TripletRow<K> tr;
    triplet_row_load<T, K>(tp, row, l, lpr, tr);
    triplet_row_finish<T, K>(tp, tr, l, lpr, reduce, loss_acc, act_acc);
}

TFC_HD void triplet_outputs(const TripletParams& tp, double sum, double act) {
    const double mean = sum / (double)tp.rows;
    tp.out[0] = (float)((double)tp.weight * mean);
    tp.out[1] = (float)mean;
    tp.out[2] = (float)(act / (double)tp.rows);
    tp.out[3] = 0.f;
}

}  // namespace tfcfft
