// regional.cuh -- regional FFT loss on the 100 x 256 "hair" and "eyes" bands.
//
// Reference: `regional_fft_loss(fake_B, real_B)`, TFC-GAN-FFT/TFCGAN_multigpu_patchFFT_withregion_FFT.py:353-402:
// rows 0..99 and 100..199 of the 256 x 256 image, each through the same grey -> rfft2 -> abs / arctan2 pipeline as
// the patch losses (-> 100 x 129 amplitude and phase), nn.L1Loss per band, the two bands SUMMED, 1/2 (amp + pha).
//
// One CTA holds one band as a complex tile z = fake + i real (100 x 256, pitch 257, 205.6 KB of the 227 KB) and runs
// the whole fused pipeline of the square kernels on it: rows are 256-point transforms (the radix-16 x 16 passes of
// fft_core.cuh on 64 + 32 + 4 lines), columns are 100-point transforms factored 4 x 25 (Cooley-Tukey: radix-4
// butterflies across the four 25-row blocks, W_100 twiddles, then a 25-point DFT per block with compile-time
// W_25 powers in registers), Hermitian un-mixing + loss + spectral gradient per half-plane bin (bin_eval), the
// inverse transforms, and the gradient store.  Frequency ky sits in row 25 (ky mod 4) + ky div 4; column
// frequencies sit at the digit-reversed positions of the 256-point passes.
#pragma once
#include "spectral_core.cuh"

namespace tfcfft {

struct RegCfg {
    static constexpr int H = 100, W = 256, LD = 257, NT = 512, BANDS = 2;
    static constexpr size_t SMEM = ((size_t)H * LD + W + 100) * sizeof(float2);  // tile + W_256 table + W_100 table
};

constexpr float kCos25[25] = {1.f,           0.968583161f,  0.87630668f,   0.728968627f,  0.535826795f,  0.309016994f,  0.0627905195f,
                              -0.187381315f, -0.425779292f, -0.63742399f,  -0.809016994f, -0.929776486f, -0.992114701f, -0.992114701f,
                              -0.929776486f, -0.809016994f, -0.63742399f,  -0.425779292f, -0.187381315f, 0.0627905195f, 0.309016994f,
                              0.535826795f,  0.728968627f,  0.87630668f,   0.968583161f};
constexpr float kSin25[25] = {0.f,           0.248689887f,  0.481753674f,  0.684547106f,  0.844327926f,  0.951056516f,  0.998026728f,
                              0.982287251f,  0.904827052f,  0.770513243f,  0.587785252f,  0.368124553f,  0.125333234f,  -0.125333234f,
                              -0.368124553f, -0.587785252f, -0.770513243f, -0.904827052f, -0.982287251f, -0.998026728f, -0.951056516f,
                              -0.844327926f, -0.684547106f, -0.481753674f, -0.248689887f};

TFC_HD constexpr int reg_row_of_freq(int ky) { return 25 * (ky & 3) + (ky >> 2); }

// out[k2] = sum_n2 u[n2] W_25^{+-n2 k2}; everything compile-time indexed
template <bool INV, int K2, int N2>
TFC_HD void dft25_term(const float2* u, float2& acc) {
    if constexpr (N2 < 25) {
        constexpr int t = (N2 * K2) % 25;
        constexpr float c = kCos25[t], sn = INV ? kSin25[t] : -kSin25[t];
        if constexpr (t == 0) {
            acc.x += u[N2].x;
            acc.y += u[N2].y;
        } else {
            acc.x = fmaf(u[N2].x, c, fmaf(-u[N2].y, sn, acc.x));
            acc.y = fmaf(u[N2].x, sn, fmaf(u[N2].y, c, acc.y));
        }
        dft25_term<INV, K2, N2 + 1>(u, acc);
    }
}
template <bool INV, int K2>
TFC_HD void dft25_all(const float2* u, float2* out) {
    if constexpr (K2 < 25) {
        float2 acc = make_float2(0.f, 0.f);
        dft25_term<INV, K2, 0>(u, acc);
        out[K2] = acc;
        dft25_all<INV, K2 + 1>(u, out);
    }
}

template <class Ctx>
TFC_HD void reg_fill_w100(const Ctx& ctx, float2* w100) {
    for (int t = ctx.tid; t < 100; t += ctx.nthreads) {
        float sn, cs;
#ifdef __CUDA_ARCH__
        sincospif(2.0f * (float)t / 100.0f, &sn, &cs);
#else
        const double a = 2.0 * 3.14159265358979323846 * (double)t / 100.0;
        sn = (float)sin(a);
        cs = (float)cos(a);
#endif
        w100[t] = make_float2(cs, -sn);
    }
}

// 100-point transforms down every column (in place).  Forward: natural rows -> frequency ky at reg_row_of_freq(ky);
// inverse: the exact reverse (unnormalised).
template <bool INV, class Ctx>
TFC_HD void reg_cols(const Ctx& ctx, float2* s, const float2* w100) {
    constexpr int LD = RegCfg::LD, W = RegCfg::W;
    if constexpr (!INV) {
        for (int it = ctx.tid; it < W * 25; it += ctx.nthreads) {  // radix-4 across the blocks + W_100^{n2 k1}
            const int x = it % W, n2 = it / W;
            float2 v[4];
#pragma unroll
            for (int n1 = 0; n1 < 4; ++n1) v[n1] = s[(n2 + 25 * n1) * LD + x];
            Dft<4, false>::run(v);
#pragma unroll
            for (int k1 = 1; k1 < 4; ++k1) v[k1] = cmul(v[k1], w100[n2 * k1]);
#pragma unroll
            for (int k1 = 0; k1 < 4; ++k1) s[(25 * k1 + n2) * LD + x] = v[k1];
        }
        ctx.sync();
    }
    for (int it = ctx.tid; it < W * 4; it += ctx.nthreads) {  // 25-point DFT of block k1
        const int x = it % W, k1 = it / W;
        float2 u[25], o[25];
#pragma unroll
        for (int n = 0; n < 25; ++n) u[n] = s[(25 * k1 + n) * LD + x];
        dft25_all<INV, 0>(u, o);
#pragma unroll
        for (int n = 0; n < 25; ++n) s[(25 * k1 + n) * LD + x] = o[n];
    }
    ctx.sync();
    if constexpr (INV) {
        for (int it = ctx.tid; it < W * 25; it += ctx.nthreads) {
            const int x = it % W, n2 = it / W;
            float2 v[4];
#pragma unroll
            for (int k1 = 0; k1 < 4; ++k1) v[k1] = s[(25 * k1 + n2) * LD + x];
#pragma unroll
            for (int k1 = 1; k1 < 4; ++k1) v[k1] = cmulc(v[k1], w100[n2 * k1]);
            Dft<4, true>::run(v);
#pragma unroll
            for (int n1 = 0; n1 < 4; ++n1) s[(n2 + 25 * n1) * LD + x] = v[n1];
        }
        ctx.sync();
    }
}

// 256-point transforms along the 100 rows: the pass code wants a power-of-two line count -> 64 + 32 + 4 lines
template <bool INV, class Ctx>
TFC_HD void reg_rows(const Ctx& ctx, float2* s, const float2* tw) {
    constexpr int LD = RegCfg::LD;
    fft_lines<256, INV>(ctx, s, 1, LD, 6, tw);
    fft_lines<256, INV>(ctx, s + 64 * LD, 1, LD, 5, tw);
    fft_lines<256, INV>(ctx, s + 96 * LD, 1, LD, 2, tw);
}

// ---- materialised band spectra (the reference's reg_fft, withregion_FFT.py:358-371) and their backward ----------
// layout [unit][100][129], unit = (n * C' + ch) * 2 + band; spec_shift = np.fft.fftshift over both axes (:253)
TFC_HD long long reg_spec_index(const Params& prm, int unit, int ky, int kx) {
    constexpr int H = RegCfg::H, WH = RegCfg::W / 2 + 1;
    int r = ky, c = kx;
    if (prm.spec_shift) {
        r = (ky + H / 2) % H;
        c = (kx + WH / 2) % WH;
    }
    return ((long long)unit * H + r) * WH + c;
}
TFC_HD void reg_bin_emit(const Params& prm, long long i, float2 zk, float2 zm) {
    const float fx = zk.x + zm.x, fy = zk.y - zm.y;  // 2F
    if (prm.spec_out[0]) prm.spec_out[0][i] = 0.5f * sqrtf(fx * fx + fy * fy);
    if (prm.spec_out[1]) prm.spec_out[1][i] = atan2f(fy, fx);
}
TFC_HD float2 reg_bin_bwd(const Params& prm, long long i, float2 zk, float2 zm) {
    const float fx = zk.x + zm.x, fy = zk.y - zm.y;
    const float f2 = sqrtf(fx * fx + fy * fy);
    const float finv = f2 > 0.f ? 1.0f / f2 : 0.f;
    const float ga = prm.spec_gin[0] ? prm.spec_gin[0][i] : 0.f;
    const float gp = prm.spec_gin[1] ? prm.spec_gin[1][i] : 0.f;
    const float ca = ga * finv, cp = gp * 2.f * finv * finv;  // d|F|/dF = F2/|F2|, d angle/dF = 2 i F2/|F2|^2
    return make_float2(ca * fx - cp * fy, ca * fy + cp * fx);
}

// loss + spectral gradient: every half-plane bin (ky, kx <= 128) is owned by exactly one item together with its mirror
template <class Ctx>
TFC_HD void reg_bins(const Ctx& ctx, const Params& prm, int unit, float2* s, float& accA, float& accP) {
    constexpr int LD = RegCfg::LD, H = RegCfg::H;
    const bool want_grad = prm.grad != nullptr;
    const int mode = prm.spec_mode;  // 0 loss, 1 emit spectra, 2 backward of the spectra
    const float2 z0 = make_float2(0.f, 0.f);
    for (int it = ctx.tid; it < H * 127; it += ctx.nthreads) {  // kx = 1..127: the mirror lies outside the half plane
        const int ky = it % H, kx = 1 + it / H;
        float2* pk = s + reg_row_of_freq(ky) * LD + pos_of_freq<256>(kx);
        float2* pm = s + reg_row_of_freq((H - ky) % H) * LD + pos_of_freq<256>(256 - kx);
        if (mode == 1) {
            reg_bin_emit(prm, reg_spec_index(prm, unit, ky, kx), *pk, *pm);
            continue;
        }
        const float2 g = mode == 2 ? reg_bin_bwd(prm, reg_spec_index(prm, unit, ky, kx), *pk, *pm) : bin_eval(prm, *pk, *pm, 1.f, accA, accP);
        if (want_grad) {
            *pk = g;
            *pm = z0;
        }
    }
    for (int it = ctx.tid; it < 2 * 51; it += ctx.nthreads) {  // self-conjugate columns kx = 0, 128: row pairs (ky, -ky)
        const int ky = it % 51, kx = (it / 51) * 128, kym = (H - ky) % H;
        float2* pk = s + reg_row_of_freq(ky) * LD + pos_of_freq<256>(kx);
        float2* pm = s + reg_row_of_freq(kym) * LD + pos_of_freq<256>(kx);
        const float2 zk = *pk, zm = *pm;
        if (mode == 1) {
            reg_bin_emit(prm, reg_spec_index(prm, unit, ky, kx), zk, zm);
            if (kym != ky) reg_bin_emit(prm, reg_spec_index(prm, unit, kym, kx), zm, zk);
            continue;
        }
        const float2 g = mode == 2 ? reg_bin_bwd(prm, reg_spec_index(prm, unit, ky, kx), zk, zm) : bin_eval(prm, zk, zm, 1.f, accA, accP);
        if (kym != ky) {
            const float2 g2 = mode == 2 ? reg_bin_bwd(prm, reg_spec_index(prm, unit, kym, kx), zm, zk) : bin_eval(prm, zm, zk, 1.f, accA, accP);
            if (want_grad) *pm = g2;
        }
        if (want_grad) *pk = g;
    }
}

// one band of one image (unit = (n * C' + ch) * 2 + band)
template <typename T, bool LUMA3, class Ctx>
TFC_HD void regional_process(const Ctx& ctx, const Params& prm, int unit, float2* s, const float2* tw, const float2* w100,
                             float& accA, float& accP) {
    constexpr int LD = RegCfg::LD, H = RegCfg::H, W = RegCfg::W, XV = W / 4;
    const int band = unit % RegCfg::BANDS, ch = (unit / RegCfg::BANDS) % prm.cprime, n = unit / (RegCfg::BANDS * prm.cprime);
    const T* fb = static_cast<const T*>(prm.fake) + n * prm.fs[0] + ch * prm.fs[1] + (long long)(band * H) * prm.fs[2];
    const T* rb = static_cast<const T*>(prm.real) + n * prm.rs[0] + ch * prm.rs[1] + (long long)(band * H) * prm.rs[2];
    for (int it = ctx.tid; it < H * XV; it += ctx.nthreads) {
        const int x = (it % XV) * 4, y = it / XV;
        float f[4], r[4];
        load_px4<T, LUMA3>(prm, fb, prm.fs, y, x, f);
        load_px4<T, LUMA3>(prm, rb, prm.rs, y, x, r);
#pragma unroll
        for (int i = 0; i < 4; ++i) s[y * LD + x + i] = make_float2(f[i], r[i]);
    }
    ctx.sync();
    reg_rows<false>(ctx, s, tw);
    reg_cols<false>(ctx, s, w100);
    reg_bins(ctx, prm, unit, s, accA, accP);
    ctx.sync();
    if (prm.grad != nullptr && prm.spec_mode != 1) {
        reg_cols<true>(ctx, s, w100);
        reg_rows<true>(ctx, s, tw);
        T* gb = static_cast<T*>(prm.grad) + n * prm.gs[0] + ch * prm.gs[1] + (long long)(band * H) * prm.gs[2];
        constexpr int NC = LUMA3 ? 3 : 1;
        const GradOut go = grad_out(prm);
        for (int it = ctx.tid; it < H * XV; it += ctx.nthreads) {
            const int x = (it % XV) * 4, y = it / XV;
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                float v[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) v[i] = go.w[c] * s[y * LD + x + i].x;
                grad_store4<T>(go, gb + (long long)y * prm.gs[2] + c * prm.gs[1] + x, v);
            }
        }
        ctx.sync();
    }
}

}  // namespace tfcfft
