// pair_tile.cuh -- the fast path for 64x64 patches (16-patch loss on 256x256, the north-star shape).
//
// Two tiles (A, B) are processed TOGETHER by one CTA: every shared-memory element is a float4
// (reA, reB, imA, imB) and every arithmetic instruction of the transforms is a packed f32x2 op
// (FADD2 / FMUL2 / FFMA2) whose two lanes belong to the two tiles.  Compared with the generic
// resident kernel (spectral_kernels.cuh) this
//   * halves the issue slots of all butterfly / twiddle / un-mixing arithmetic,
//   * makes every shared-memory access a conflict-free 128-bit LDS/STS,
//   * enumerates exactly the half-plane bins (no idle lanes in the loss pass),
//   * runs the inverse column transforms on the 33 non-zero columns only,
//   * replaces atan2f / sqrtf / division by MUFU approximations + a packed minimax polynomial.
// The stage functions are __host__ __device__ so the CPU emulation (emu.cu) executes the same code.
#pragma once
#include "spectral_core.cuh"

namespace tfcfft {

template <int P>
struct PairCfg {
    static constexpr int LD = P + 1;          // float4 row pitch, odd: row and column walks are conflict-free
    static constexpr int NT_COMPUTE = 512;    // 16 warps transform the current pair
    static constexpr int NT_LOAD = 256;       // 8 warps stream the next pair from HBM into the other buffer
    static constexpr int NT = NT_COMPUTE + NT_LOAD;
    static constexpr size_t TILE_BYTES = (size_t)P * LD * sizeof(float4);
    static constexpr size_t SMEM = 2 * TILE_BYTES + 2 * (size_t)P * sizeof(float4);  // two work buffers + 2 twiddle tables
};

// ---- MUFU-level approximations (1-2 ulp), exact libm on the host emulation -------------------
TFC_HD float fast_rcp(float x) {
#ifdef __CUDA_ARCH__
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    return 1.0f / x;
#endif
}
TFC_HD float fast_rsqrt(float x) {
#ifdef __CUDA_ARCH__
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    return 1.0f / sqrtf(x);
#endif
}
TFC_HD float fast_sqrt(float x) {
#ifdef __CUDA_ARCH__
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    return sqrtf(x);
#endif
}

// atan2 of two lanes.  atan(t) = t * P7(t^2) on [0,1] (minimax, |err| < 4e-8 in exact arithmetic,
// ~1.5e-7 evaluated in fp32); octant fix-ups per lane.  atan2(+-0, 0) = 0 and atan2(+0, x<0) = +pi as
// NumPy / torch.
TFC_HD float2 atan2_pair(float2 y, float2 x) {
    const float ax0 = fabsf(x.x), ay0 = fabsf(y.x), ax1 = fabsf(x.y), ay1 = fabsf(y.y);
    const float mx0 = fmaxf(ax0, ay0), mn0 = fminf(ax0, ay0);
    const float mx1 = fmaxf(ax1, ay1), mn1 = fminf(ax1, ay1);
    float2 t;
    t.x = mx0 > 1e-30f ? mn0 * fast_rcp(mx0) : 0.f;
    t.y = mx1 > 1e-30f ? mn1 * fast_rcp(mx1) : 0.f;
    const float2 s = p_mul(t, t);
    float2 p = p_dup(-4.0545672114e-03f);
    p = p_fma(p, s, p_dup(2.1862957868e-02f));
    p = p_fma(p, s, p_dup(-5.5912326758e-02f));
    p = p_fma(p, s, p_dup(9.6421973272e-02f));
    p = p_fma(p, s, p_dup(-1.3908629550e-01f));
    p = p_fma(p, s, p_dup(1.9946565651e-01f));
    p = p_fma(p, s, p_dup(-3.3329860784e-01f));
    p = p_fma(p, s, p_dup(9.9999933558e-01f));
    float2 r = p_mul(p, t);
    constexpr float kPi = 3.14159265358979f, kHalfPi = 1.57079632679490f;
    if (ay0 > ax0) r.x = kHalfPi - r.x;
    if (ay1 > ax1) r.y = kHalfPi - r.y;
    if (x.x < 0.f) r.x = kPi - r.x;
    if (x.y < 0.f) r.y = kPi - r.y;
    if (y.x < 0.f) r.x = -r.x;
    if (y.y < 0.f) r.y = -r.y;
    return r;
}

// unsigned angle of two lanes: atan2(|s|, c) in [0, pi] (same polynomial as atan2_pair, no final sign fix-up)
TFC_HD float2 atan2_abs_pair(float2 sn, float2 cs) {
    const float ax0 = fabsf(cs.x), ay0 = fabsf(sn.x), ax1 = fabsf(cs.y), ay1 = fabsf(sn.y);
    const float mx0 = fmaxf(ax0, ay0), mn0 = fminf(ax0, ay0);
    const float mx1 = fmaxf(ax1, ay1), mn1 = fminf(ax1, ay1);
    float2 t;
    t.x = mx0 > 1e-30f ? mn0 * fast_rcp(mx0) : 0.f;
    t.y = mx1 > 1e-30f ? mn1 * fast_rcp(mx1) : 0.f;
    const float2 s = p_mul(t, t);
    float2 p = p_dup(-4.0545672114e-03f);
    p = p_fma(p, s, p_dup(2.1862957868e-02f));
    p = p_fma(p, s, p_dup(-5.5912326758e-02f));
    p = p_fma(p, s, p_dup(9.6421973272e-02f));
    p = p_fma(p, s, p_dup(-1.3908629550e-01f));
    p = p_fma(p, s, p_dup(1.9946565651e-01f));
    p = p_fma(p, s, p_dup(-3.3329860784e-01f));
    p = p_fma(p, s, p_dup(9.9999933558e-01f));
    float2 r = p_mul(p, t);
    constexpr float kPi = 3.14159265358979f, kHalfPi = 1.57079632679490f;
    if (ay0 > ax0) r.x = kHalfPi - r.x;
    if (ay1 > ax1) r.y = kHalfPi - r.y;
    if (cs.x < 0.f) r.x = kPi - r.x;
    if (cs.y < 0.f) r.y = kPi - r.y;
    return r;
}

TFC_HD int f_bits(float f) {
#ifdef __CUDA_ARCH__
    return __float_as_int(f);
#else
    int i;
    memcpy(&i, &f, 4);
    return i;
#endif
}
TFC_HD float bits_f(int i) {
#ifdef __CUDA_ARCH__
    return __int_as_float(i);
#else
    float f;
    memcpy(&f, &i, 4);
    return f;
#endif
}

// Phase difference d = angle(F) - angle(R) in (-2 pi, 2 pi) -- the reference subtracts two arctan2 values without
// re-wrapping -- from ONE arctangent: a = atan2(|Im(F conj R)|, Re(F conj R)) is the unsigned angle between F and R,
// and the half planes of F and R (the sign bits of their imaginary parts) say which multiple of 2 pi separates d
// from the wrapped angle:
//   same half plane:       d = sign(Im(F conj R)) * a
//   F upper, R lower:      d in (0, 2 pi):   d = a if Im(F conj R) >= 0 else 2 pi - a          (and mirrored)
// The decisions only read sign bits; where such a sign is decided by cancellation (d near 0 / +-pi in the same half
// plane, d near pi across) the two candidate values coincide, and at the real axis the signed zeros of the fused
// multiply-add give the IEEE answers (atan2(+0, x < 0) = +pi).  Returns |d| in `absd` and the sign bit of d in the
// top bit of `sgn` (per lane).
TFC_HD void phase_delta(float fx, float fy, float rx, float ry, float sn, float a, float& absd, int& sgn) {
    constexpr float kTwoPi = 6.28318530717958648f;
    const int bf = f_bits(fy), br = f_bits(ry), bs = f_bits(sn);
    const int opp = bf ^ br;              // top bit: F and R in different half planes
    sgn = opp < 0 ? bf : bs;
    const int wrap = (bs ^ bf) & opp;     // top bit: different half planes and the short way round is the wrong way
    absd = wrap < 0 ? kTwoPi - a : a;
    (void)fx;
    (void)rx;
}

TFC_HD float sign_of(float d) {  // copysign(1, d): the sign(0) = 0 case only matters when F == 0, where the
                                 // gradient is already zeroed through 1/|F| := 0
#if defined(__CUDA_ARCH__)
    return __int_as_float((__float_as_int(d) & 0x80000000) | 0x3f800000);
#else
    return d < 0.f ? -1.f : 1.f;
#endif
}

// Loss terms and spectral gradient of one half-plane bin for both tiles.  zk = Z(k), zm = Z(-k).
TFC_HD c2 bin_eval_pair(const Params& prm, bool mse, bool phase, c2 zk, c2 zm, float2& accA, float2& accP) {
    const float2 fx = p_add(zk.re, zm.re), fy = p_sub(zk.im, zm.im);  // 2F
    const float2 rx = p_add(zk.im, zm.im), ry = p_sub(zm.re, zk.re);  // 2R
    const float2 fsq = p_fma(fx, fx, p_mul(fy, fy)), rsq = p_fma(rx, rx, p_mul(ry, ry));
    float2 finv, f2, r2;
    finv.x = fsq.x > 1e-35f ? fast_rsqrt(fsq.x) : 0.f;
    finv.y = fsq.y > 1e-35f ? fast_rsqrt(fsq.y) : 0.f;
    f2 = p_mul(fsq, finv);  // |2F|
    r2.x = fast_sqrt(rsq.x);
    r2.y = fast_sqrt(rsq.y);
    const float2 da = p_mul(p_sub(f2, r2), p_dup(0.5f));
    float2 ga;
    if (mse) {
        accA = p_fma(da, da, accA);
        ga = p_add(da, da);
    } else {
        accA.x += fabsf(da.x);
        accA.y += fabsf(da.y);
        ga = make_float2(sign_of(da.x), sign_of(da.y));
    }
    const float2 ca = p_mul(p_mul(ga, finv), p_dup(prm.sa));
    c2 g = make_c2(p_mul(ca, fx), p_mul(ca, fy));
    if (phase) {
        const float2 cs = p_fma(fx, rx, p_mul(fy, ry));          // Re(F conj R)
        const float2 sn = p_fma(fy, rx, p_neg(p_mul(fx, ry)));   // Im(F conj R)
        const float2 a = atan2_abs_pair(sn, cs);
        float2 ad;
        int s0, s1;
        phase_delta(fx.x, fy.x, rx.x, ry.x, sn.x, a.x, ad.x, s0);
        phase_delta(fx.y, fy.y, rx.y, ry.y, sn.y, a.y, ad.y, s1);
        float2 gp;
        if (mse) {
            accP = p_fma(ad, ad, accP);
            gp = make_float2(bits_f((s0 & 0x80000000) | (f_bits(ad.x + ad.x) & 0x7fffffff)),
                             bits_f((s1 & 0x80000000) | (f_bits(ad.y + ad.y) & 0x7fffffff)));
        } else {
            accP = p_add(accP, ad);
            gp = make_float2(bits_f((s0 & 0x80000000) | 0x3f800000), bits_f((s1 & 0x80000000) | 0x3f800000));
        }
        const float2 cp = p_mul(p_mul(gp, p_mul(finv, finv)), p_dup(2.f * prm.sp));
        g.re = p_fma(p_neg(cp), fy, g.re);
        g.im = p_fma(cp, fx, g.im);
    }
    return g;
}

// XOR swizzle of the freshly loaded tile: the loader's 4-pixel STS.128 bursts become conflict-free.
TFC_HD int swz(int x) { return x ^ ((x >> 3) & 3); }

// (wr, wr, wi, wi) table of e^{-2 pi i t / P}
template <int P, class Ctx>
TFC_HD void fill_twiddles4(const Ctx& ctx, float4* tw) {
    for (int t = ctx.tid; t < P; t += ctx.nthreads) {
        float sn, cs;
#ifdef __CUDA_ARCH__
        sincospif(2.0f * (float)t / (float)P, &sn, &cs);
#else
        const double a = 2.0 * 3.14159265358979323846 * (double)t / (double)P;
        sn = (float)sin(a);
        cs = (float)cos(a);
#endif
        tw[t] = make_float4(cs, cs, -sn, -sn);
    }
}

// Row-pass copy of the twiddles, laid out [k][j] = W_P^{j k}: the M lanes of a row read consecutive float4
// (the plain table would be read with stride k: bank conflicts for even k).
template <int P, class Ctx>
TFC_HD void fill_row_twiddles4(const Ctx& ctx, const float4* tw, float4* twr) {
    constexpr int R = Plan<P>::R1, M = P / R;
    for (int t = ctx.tid; t < R * M; t += ctx.nthreads) twr[t] = tw[((t / M) * (t % M)) % P];
}

// ---- stage 0: global -> luma -> packed tile pair ----------------------------------------------
// Base pointers of the four source tiles (fake A, real A, fake B, real B), computed once per pair.
template <typename T>
struct PairSrc {
    const T* p[4];
    int sh[4], sc[4];  // row / channel strides in elements (tile-local offsets fit 32 bits)
};
template <int P, typename T>
TFC_HD PairSrc<T> pair_src(const Params& prm, const TileCoord& ta, const TileCoord& tb) {
    PairSrc<T> r;
    r.p[0] = tile_ptr<T>(prm.fake, prm.fs, ta, P);
    r.p[1] = real_tile_ptr<T>(prm, ta, P);
    r.p[2] = tile_ptr<T>(prm.fake, prm.fs, tb, P);
    r.p[3] = real_tile_ptr<T>(prm, tb, P);
    r.sh[0] = r.sh[2] = (int)prm.fs[2];
    r.sh[1] = r.sh[3] = (int)prm.rs[2];
    r.sc[0] = r.sc[2] = (int)prm.fs[1];
    r.sc[1] = r.sc[3] = (int)prm.rs[1];
    return r;
}

template <int P, typename T, bool LUMA3, class Ctx>
TFC_HD void pair_load(const Ctx& ctx, const Params& prm, const TileCoord& ta, const TileCoord& tb, float4* s) {
    constexpr int LD = PairCfg<P>::LD, XV = P / 4, NC = LUMA3 ? 3 : 1;
    const PairSrc<T> src = pair_src<P, T>(prm, ta, tb);
    const bool quant = (prm.flags & TFCFFT_QUANTIZE_U8) != 0;
    const float2 w0 = p_dup(prm.lw[0]), w1 = p_dup(prm.lw[1]), w2 = p_dup(prm.lw[2]);
    for (int it = ctx.tid; it < P * XV; it += ctx.nthreads) {
        const int x = (it % XV) * 4, y = it / XV;
        float4* row = s + y * LD;
        float2 v[2][4];  // [fake | real][pixel] = (tile A, tile B)
        // two half-items (fake A+B, then real A+B): 2*NC 128-bit loads in flight per thread each
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            float raw[2][NC][4];
#pragma unroll
            for (int ab = 0; ab < 2; ++ab)
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    const int w = h + 2 * ab;
                    IO<T>::load4(src.p[w] + y * src.sh[w] + c * src.sc[w] + x, raw[ab][c]);
                }
            if (!quant) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float2 f = p_mul(w0, make_float2(raw[0][0][i], raw[1][0][i]));
                    if constexpr (LUMA3) {
                        f = p_fma(w1, make_float2(raw[0][1][i], raw[1][1][i]), f);
                        f = p_fma(w2, make_float2(raw[0][2][i], raw[1][2][i]), f);
                    }
                    v[h][i] = f;
                }
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float q[2];
#pragma unroll
                    for (int ab = 0; ab < 2; ++ab) {
                        if constexpr (LUMA3)
                            q[ab] = (float)((19595 * IO<T>::quant(raw[ab][0][i]) + 38470 * IO<T>::quant(raw[ab][1][i]) +
                                             7471 * IO<T>::quant(raw[ab][2][i]) + 0x8000) >> 16);
                        else
                            q[ab] = (float)IO<T>::quant(raw[ab][0][i]);
                    }
                    v[h][i] = make_float2(q[0], q[1]);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) row[swz(x + i)] = make_float4(v[0][i].x, v[0][i].y, v[1][i].x, v[1][i].y);
    }
}

// ---- stages 1+2: forward row passes.  The first works in place on the swizzled addresses (each
// task rewrites exactly the words it read); the second owns whole groups of R2 positions, reads them
// through the swizzle and writes plain positions, which removes the swizzle for free. ------------
template <int P, class Ctx>
TFC_HD void pair_rows_first(const Ctx& ctx, float4* s, const float4* twr) {
    constexpr int R = Plan<P>::R1, M = P / R, LD = PairCfg<P>::LD;
    for (int t = ctx.tid; t < P * M; t += ctx.nthreads) {
        const int j = t % M, y = t / M;  // consecutive threads: consecutive columns of one row
        float4* row = s + y * LD;
        c2 v[R];
#pragma unroll
        for (int m = 0; m < R; ++m) v[m] = Cx<float4>::ld(row[swz(j + m * M)]);
        Dft<R, false>::run(v);
#pragma unroll
        for (int k = 1; k < R; ++k) v[k] = cmul(v[k], twr[k * M + j]);
#pragma unroll
        for (int k = 0; k < R; ++k) row[swz(j + k * M)] = Cx<float4>::st(v[k]);
    }
}

template <int P, class Ctx>
TFC_HD void pair_rows_second(const Ctx& ctx, float4* s) {
    constexpr int R = Plan<P>::R2, NB = P / R, LD = PairCfg<P>::LD;
    for (int t = ctx.tid; t < P * NB; t += ctx.nthreads) {
        const int y = t % P, k1 = t / P;  // consecutive threads: consecutive rows (odd pitch: conflict-free)
        float4* grp = s + y * LD + k1 * R;
        c2 v[R];
#pragma unroll
        for (int j = 0; j < R; ++j) v[j] = Cx<float4>::ld(s[y * LD + swz(k1 * R + j)]);
        Dft<R, false>::run(v);
#pragma unroll
        for (int k = 0; k < R; ++k) grp[k] = Cx<float4>::st(v[k]);
    }
}

// ---- loss pass over exactly the half-plane bins -------------------------------------------------
// P*P/2 work items: P*(P/2-1) regular bins (columns kx = 1..P/2-1, one item per bin; the mirror position
// is zeroed for the inverse row passes) followed by P items for the two self-conjugate columns kx = 0 and
// kx = P/2 (an item owns rows ky and -ky; the four self-conjugate bins of a column pair up as one item).
template <int P, class Ctx>
TFC_HD void pair_bins(const Ctx& ctx, const Params& prm, float4* s, float2& accA, float2& accP) {
    constexpr int LD = PairCfg<P>::LD, H = P / 2, NREG = P * (H - 1);
    const bool mse = (prm.flags & TFCFFT_DIST_MSE) != 0, phase = !(prm.flags & TFCFFT_NO_PHASE);
    const bool want_grad = prm.grad != nullptr;
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    // two items per iteration: their (long, serial) dependency chains interleave
    for (int it0 = ctx.tid; it0 < P * H; it0 += 2 * ctx.nthreads) {
        float4* pk[2];
        float4* pm[2];
        c2 zk[2], zm[2];
        bool live[2], reg[2], selfpair[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int it = it0 + u * ctx.nthreads;
            live[u] = it < P * H;
            reg[u] = it < NREG;
            selfpair[u] = false;
            int qy, qx, qym, qxm;
            if (reg[u] || !live[u]) {
                const int iq = live[u] ? it : 0;
                qy = iq % P;  // consecutive threads: consecutive row positions
                const int kx = 1 + iq / P;
                qx = pos_of_freq<P>(kx);
                qxm = pos_of_freq<P>(P - kx);
                qym = neg_pos<P>(qy);
            } else {
                const int sp = it - NREG, r = sp % H;
                qx = qxm = pos_of_freq<P>((sp / H) * H);
                // r == 0: the two self-conjugate bins ky = 0 and ky = P/2; else the pair (ky, -ky) = (r, P - r)
                qy = pos_of_freq<P>(r);
                qym = pos_of_freq<P>(r == 0 ? H : P - r);
                selfpair[u] = (r == 0);
            }
            pk[u] = s + qy * LD + qx;
            pm[u] = s + qym * LD + qxm;
            zk[u] = Cx<float4>::ld(*pk[u]);
            zm[u] = Cx<float4>::ld(*pm[u]);
        }
        c2 g[2], g2[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            if (!live[u]) continue;
            if (reg[u]) {
                g[u] = bin_eval_pair(prm, mse, phase, zk[u], zm[u], accA, accP);
            } else {
                g[u] = bin_eval_pair(prm, mse, phase, zk[u], selfpair[u] ? zk[u] : zm[u], accA, accP);
                g2[u] = bin_eval_pair(prm, mse, phase, zm[u], selfpair[u] ? zm[u] : zk[u], accA, accP);
            }
        }
        if (want_grad) {
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                if (!live[u]) continue;
                *pk[u] = Cx<float4>::st(g[u]);
                *pm[u] = reg[u] ? zero : Cx<float4>::st(g2[u]);
            }
        }
    }
}

// ---- inverse column passes on the P/2+1 non-zero columns only ----------------------------------
// column positions with kx <= P/2: {k1*R2 + k2 : k2 < R2/2} (P/2 of them) plus the Nyquist column R2/2
template <int P, int R, int L, class Ctx>
TFC_HD void pair_cols_inv_pass(const Ctx& ctx, float4* s, const float4* tw) {
    constexpr int LD = PairCfg<P>::LD, R2 = Plan<P>::R2, HR = R2 / 2, M = L / R, JT = P / R;
    constexpr int NL = P / 2 + 1;
    for (int t = ctx.tid; t < NL * JT; t += ctx.nthreads) {
        const int li = t % NL, jj = t / NL;
        const int qx = li < P / 2 ? (li / HR) * R2 + (li % HR) : HR;
        const int blk = jj / M, j = jj % M;
        float4* base = s + (blk * L + j) * LD + qx;
        c2 v[R];
#pragma unroll
        for (int m = 0; m < R; ++m) v[m] = Cx<float4>::ld(base[m * M * LD]);
        if constexpr (M > 1) {
#pragma unroll
            for (int k = 1; k < R; ++k) v[k] = cmulc(v[k], tw[(P / L) * j * k]);
        }
        Dft<R, true>::run(v);
#pragma unroll
        for (int k = 0; k < R; ++k) base[k * M * LD] = Cx<float4>::st(v[k]);
    }
}

// ---- last inverse row pass: the real parts (the two gradient tiles) stay in shared memory ---------
// Output pixel x of row y goes to float4 slot swz(x) as (gA, gB, -, -).  The eight tasks of a row sit in
// eight adjacent lanes of one warp; a warp-level barrier separates their reads from the swizzled writes.
template <int P, class Ctx>
TFC_HD void pair_rows_last(const Ctx& ctx, float4* s, const float4* twr) {
    constexpr int R = Plan<P>::R1, M = P / R, LD = PairCfg<P>::LD;
    static_assert(M <= 32 && (32 % M) == 0, "the tasks of one row must share a warp");
    for (int t = ctx.tid; t < P * M; t += ctx.nthreads) {
        const int j = t % M, y = t / M;
        float4* row = s + y * LD;
#ifndef __CUDA_ARCH__
        // serial emulation: tasks of a row run one after the other, so they read a snapshot of the row
        static thread_local float4 snap[P];
        if (j == 0)
            for (int q = 0; q < P; ++q) snap[q] = row[q];
        const float4* src = snap;
#else
        const float4* src = row;
#endif
        c2 v[R];
#pragma unroll
        for (int k = 0; k < R; ++k) v[k] = Cx<float4>::ld(src[j + k * M]);
#pragma unroll
        for (int k = 1; k < R; ++k) v[k] = cmulc(v[k], twr[k * M + j]);
        Dft<R, true>::run(v);
        ctx.warp_sync();
#pragma unroll
        for (int m = 0; m < R; ++m) {
            float2* dst = reinterpret_cast<float2*>(row + swz(j + m * M));
            *dst = v[m].re;
        }
    }
}

// ---- gradient store: shared memory -> global, coalesced 128-bit stores -----------------------------
// Runs on the loader warps of the kernel (off the transform's critical path).
template <int P, typename T, bool LUMA3, class Ctx>
TFC_HD void pair_store(const Ctx& ctx, const Params& prm, const TileCoord& ta, const TileCoord& tb, bool b_valid,
                       const float4* s) {
    constexpr int LD = PairCfg<P>::LD, XV = P / 4, NC = LUMA3 ? 3 : 1;
    T* ga = const_cast<T*>(tile_ptr<T>(prm.grad, prm.gs, ta, P));
    T* gb = const_cast<T*>(tile_ptr<T>(prm.grad, prm.gs, tb, P));
    const int sh = (int)prm.gs[2], sc = (int)prm.gs[1];
    const GradOut go = grad_out(prm);
    for (int it = ctx.tid; it < P * XV; it += ctx.nthreads) {
        const int x = (it % XV) * 4, y = it / XV;
        const float4* row = s + y * LD;
        float a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 g = *reinterpret_cast<const float2*>(row + swz(x + i));
            a[i] = g.x;
            b[i] = g.y;
        }
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            float va[4], vb[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                va[i] = go.w[c] * a[i];
                vb[i] = go.w[c] * b[i];
            }
            grad_store4<T>(go, ga + y * sh + c * sc + x, va);
            if (b_valid) grad_store4<T>(go, gb + y * sh + c * sc + x, vb);
        }
    }
}

// ---- one tile pair: everything after the load stage -------------------------------------------
template <int P, typename T, bool LUMA3, class Ctx>
TFC_HD void pair_compute(const Ctx& ctx, const Params& prm, const TileCoord& ta, const TileCoord& tb, bool b_valid,
                         float4* s, const float4* tw, float2& accA, float2& accP) {
    using Pl = Plan<P>;
    constexpr int LD = PairCfg<P>::LD, LP = ilog2_c(P), L2 = P / Pl::R1;
    static_assert(Pl::R3 == 1, "pair path supports two-pass plans");
    ctx.mark(1);
    pair_rows_first<P>(ctx, s, tw + P);
    ctx.sync();
    ctx.mark(2);
    pair_rows_second<P>(ctx, s);
    ctx.sync();
    ctx.mark(3);
    fft_pass<P, Pl::R1, P, false>(ctx, s, LD, 1, LP, tw);   // columns: thread-fast = column
    ctx.sync();
    ctx.mark(4);
    fft_pass<P, Pl::R2, L2, false>(ctx, s, LD, 1, LP, tw);
    ctx.sync();
    ctx.mark(5);
    pair_bins<P>(ctx, prm, s, accA, accP);
    ctx.sync();
    ctx.mark(6);
    if (prm.grad != nullptr) {
        pair_cols_inv_pass<P, Pl::R2, L2>(ctx, s, tw);
        ctx.sync();
        ctx.mark(7);
        pair_cols_inv_pass<P, Pl::R1, P>(ctx, s, tw);
        ctx.sync();
        ctx.mark(8);
        fft_pass<P, Pl::R2, L2, true>(ctx, s, 1, LD, LP, tw);
        ctx.sync();
        ctx.mark(9);
        pair_rows_last<P>(ctx, s, tw + P);
        ctx.sync();
        ctx.mark(10);
    }
}

// load + compute by the same threads (CPU emulation; the kernel splits the two across warp roles)
template <int P, typename T, bool LUMA3, class Ctx>
TFC_HD void pair_process(const Ctx& ctx, const Params& prm, int tile_a, int tile_b, bool b_valid, float4* s,
                         const float4* tw, float2& accA, float2& accP) {
    const TileCoord ta = decode_tile(prm, tile_a), tb = decode_tile(prm, tile_b);
    ctx.mark(0);
    pair_load<P, T, LUMA3>(ctx, prm, ta, tb, s);
    ctx.sync();
    pair_compute<P, T, LUMA3>(ctx, prm, ta, tb, b_valid, s, tw, accA, accP);
    if (prm.grad != nullptr) pair_store<P, T, LUMA3>(ctx, prm, ta, tb, b_valid, s);
    ctx.sync();
}


}  // namespace tfcfft
