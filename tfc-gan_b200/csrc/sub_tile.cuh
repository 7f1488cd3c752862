// sub_tile.cuh -- 128 x 128 and 256 x 256 tiles on top of the packed 64 x 64 machinery.
//
// A P x P tile (P = 64 D, D = 2 or 4) is decimated in both dimensions into D x D interleaved sub-images
// s_pq[a, b] = x[D a + p, D b + q].  The 2-D DFT factors exactly (decimation in time):
//     Z[ky' + 64 al, kx' + 64 be] = sum_{p,q} W_D^{p al + q be} * W_P^{p ky' + q kx'} * S_pq[ky', kx']
// so the heavy work -- D^2 independent 64 x 64 complex transforms -- runs in the packed, warp-specialised pair
// kernel (two sub-images with adjacent pixels q = 2i, 2i+1 are the two f32x2 lanes), and the cross-sub-image
// step is a D x D butterfly per frequency position done in registers by `combine_kernel`, which also meets
// Z(k) with Z(-k), evaluates the loss and the spectral gradient, and applies the inverse butterfly.  The three
// launches exchange the sub-spectra through an L2-sized workspace chunk (P*P*8 bytes per tile):
//   pair_kernel(mode 1): sub-images -> S_pq          (HBM read of fake / real)
//   combine_kernel<D>:   S_pq -> Z -> loss, G -> H_pq (L2 only)
//   pair_kernel(mode 2): H_pq -> gradient sub-images  (HBM write of grad)
#pragma once
#include "pair_tile.cuh"

namespace tfcfft {

struct SubUnit {
    int tile_local, plane, p, i;  // tile within the chunk, pair plane, sub-image row phase, lane-pair index
};
TFC_HD SubUnit sub_unit(int u, int d) {
    const int npp = d * d / 2, hd = d / 2;
    SubUnit r;
    r.tile_local = u / npp;
    r.plane = u % npp;
    r.p = r.plane / hd;
    r.i = r.plane % hd;
    return r;
}
TFC_HD float4* sub_plane(const Params& prm, const SubUnit& su) {
    return reinterpret_cast<float4*>(prm.zws) + ((long long)su.tile_local * (prm.sub_d * prm.sub_d / 2) + su.plane) * 4096;
}

// ---- sub-image pair -> packed work tile (same swizzled layout as pair_load) ----------------------
template <typename T, bool LUMA3, class Ctx>
TFC_HD void sub_load(const Ctx& ctx, const Params& prm, const TileCoord& tc, const SubUnit& su, float4* s) {
    constexpr int LD = 65, NC = LUMA3 ? 3 : 1;
    const int D = prm.sub_d, P = 64 * D;
    const T* fp = tile_ptr<T>(prm.fake, prm.fs, tc, P);
    const T* rp = tile_ptr<T>(prm.real, prm.rs, tc, P);
    const int fsh = (int)prm.fs[2], fsc = (int)prm.fs[1], rsh = (int)prm.rs[2], rsc = (int)prm.rs[1];
    const bool quant = (prm.flags & TFCFFT_QUANTIZE_U8) != 0;
    // one sub-image column b per lane (consecutive lanes -> consecutive 8-byte pairs of the source row), four
    // rows a, a+16, a+32, a+48 per thread in flight
    for (int it = ctx.tid; it < 64 * 16; it += ctx.nthreads) {
        const int b = it & 63, a0 = it >> 6;
        const int x = D * b + 2 * su.i;
        float raw[4][2][NC][2];  // [row][fake|real][channel][lane]
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int y = D * (a0 + 16 * r) + su.p;
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                IO<T>::load2(fp + y * fsh + c * fsc + x, raw[r][0][c]);
                IO<T>::load2(rp + y * rsh + c * rsc + x, raw[r][1][c]);
            }
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            float2 v[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if (!quant) {
                    float2 f = p_mul(p_dup(prm.lw[0]), make_float2(raw[r][h][0][0], raw[r][h][0][1]));
                    if constexpr (LUMA3) {
                        f = p_fma(p_dup(prm.lw[1]), make_float2(raw[r][h][1][0], raw[r][h][1][1]), f);
                        f = p_fma(p_dup(prm.lw[2]), make_float2(raw[r][h][2][0], raw[r][h][2][1]), f);
                    }
                    v[h] = f;
                } else {
                    float q[2];
#pragma unroll
                    for (int l = 0; l < 2; ++l) {
                        if constexpr (LUMA3)
                            q[l] = (float)((19595 * IO<T>::quant(raw[r][h][0][l]) + 38470 * IO<T>::quant(raw[r][h][1][l]) +
                                            7471 * IO<T>::quant(raw[r][h][2][l]) + 0x8000) >> 16);
                        else
                            q[l] = (float)IO<T>::quant(raw[r][h][0][l]);
                    }
                    v[h] = make_float2(q[0], q[1]);
                }
            }
            s[(a0 + 16 * r) * LD + swz(b)] = make_float4(v[0].x, v[0].y, v[1].x, v[1].y);
        }
    }
}

// ---- gradient sub-image pair (staged at the swizzled slots by pair_rows_last) -> global ------------
template <typename T, bool LUMA3, class Ctx>
TFC_HD void sub_store(const Ctx& ctx, const Params& prm, const TileCoord& tc, const SubUnit& su, const float4* s) {
    constexpr int LD = 65, NC = LUMA3 ? 3 : 1;
    const int D = prm.sub_d, P = 64 * D;
    T* gp = const_cast<T*>(tile_ptr<T>(prm.grad, prm.gs, tc, P));
    const int sh = (int)prm.gs[2], sc = (int)prm.gs[1];
    for (int it = ctx.tid; it < 64 * 16; it += ctx.nthreads) {
        const int b = it & 63, a0 = it >> 6;  // consecutive lanes: consecutive 8-byte pairs of the gradient row
        const int x = D * b + 2 * su.i;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int a = a0 + 16 * r, y = D * a + su.p;
            const float2 g = *reinterpret_cast<const float2*>(s + a * LD + swz(b));
#pragma unroll
            for (int c = 0; c < NC; ++c) IO<T>::store2(gp + y * sh + c * sc + x, prm.gw[c] * g.x, prm.gw[c] * g.y);
        }
    }
}

// ---- spectrum plane <-> work tile (plain pitch-65 layout, positions as the pair passes leave them) --
template <class Ctx>
TFC_HD void spec_store(const Ctx& ctx, const float4* s, float4* plane) {
    for (int it = ctx.tid; it < 4096; it += ctx.nthreads) plane[it] = s[(it >> 6) * 65 + (it & 63)];
}
template <class Ctx>
TFC_HD void spec_load(const Ctx& ctx, const float4* plane, float4* s) {
    for (int it = ctx.tid; it < 4096; it += ctx.nthreads) s[(it >> 6) * 65 + (it & 63)] = plane[it];
}

// forward transform of a loaded pair: rows then columns (the loss pass happens in combine_kernel)
template <class Ctx>
TFC_HD void sub_compute_fwd(const Ctx& ctx, float4* s, const float4* tw) {
    using Pl = Plan<64>;
    pair_rows_first<64>(ctx, s, tw + 64);
    ctx.sync();
    pair_rows_second<64>(ctx, s);
    ctx.sync();
    fft_pass<64, Pl::R1, 64, false>(ctx, s, 65, 1, 6, tw);
    ctx.sync();
    fft_pass<64, Pl::R2, 8, false>(ctx, s, 65, 1, 6, tw);
    ctx.sync();
}
// inverse transform of a pair of H_pq planes: all 64 columns, then rows; real parts staged for the store
template <class Ctx>
TFC_HD void sub_compute_inv(const Ctx& ctx, float4* s, const float4* tw) {
    using Pl = Plan<64>;
    fft_pass<64, Pl::R2, 8, true>(ctx, s, 65, 1, 6, tw);
    ctx.sync();
    fft_pass<64, Pl::R1, 64, true>(ctx, s, 65, 1, 6, tw);
    ctx.sync();
    fft_pass<64, Pl::R2, 8, true>(ctx, s, 1, 65, 6, tw);
    ctx.sync();
    pair_rows_last<64>(ctx, s, tw + 64);
    ctx.sync();
}

// ---- D-point butterflies over small register arrays ------------------------------------------------
TFC_HD float2 cis_neg(float frac) {  // e^{-2 pi i frac}
    float sn, cs;
#ifdef __CUDA_ARCH__
    sincospif(2.0f * frac, &sn, &cs);
#else
    const double a = 2.0 * 3.14159265358979323846 * (double)frac;
    sn = (float)sin(a);
    cs = (float)cos(a);
#endif
    return make_float2(cs, -sn);
}

// ---- D x D butterflies on two positions at once (lane x = position A, lane y = its partner B) ----
// forward: S[p][q] (sub-spectra) -> Z[al][be] (full-size spectrum entries); inverse: the unnormalised adjoint
TFC_HD c2 cmul2(c2 a, c2 w) {  // per-lane a * w
    return make_c2(p_fma(a.re, w.re, p_neg(p_mul(a.im, w.im))), p_fma(a.re, w.im, p_mul(a.im, w.re)));
}
TFC_HD c2 cmulc2(c2 a, c2 w) {  // per-lane a * conj(w)
    return make_c2(p_fma(a.re, w.re, p_mul(a.im, w.im)), p_fma(a.im, w.re, p_neg(p_mul(a.re, w.im))));
}
TFC_HD c2 cis_neg2(float fa, float fb) {
    const float2 a = cis_neg(fa), b = cis_neg(fb);
    return make_c2(make_float2(a.x, b.x), make_float2(a.y, b.y));
}
template <int D>
TFC_HD void combine_fwd2(c2 (&v)[D][D], int kyA, int kxA, int kyB, int kxB) {
    constexpr int P = 64 * D;
    const c2 wy = cis_neg2((float)kyA / (float)P, (float)kyB / (float)P);
    const c2 wx = cis_neg2((float)kxA / (float)P, (float)kxB / (float)P);
#pragma unroll
    for (int p = 0; p < D; ++p) {
        c2 w = wx;
#pragma unroll
        for (int q = 1; q < D; ++q) {
            v[p][q] = cmul2(v[p][q], w);
            if (q + 1 < D) w = cmul2(w, wx);
        }
        Dft<D, false>::run(v[p]);
    }
#pragma unroll
    for (int be = 0; be < D; ++be) {
        c2 col[D];
        c2 w = wy;
#pragma unroll
        for (int p = 0; p < D; ++p) col[p] = v[p][be];
#pragma unroll
        for (int p = 1; p < D; ++p) {
            col[p] = cmul2(col[p], w);
            if (p + 1 < D) w = cmul2(w, wy);
        }
        Dft<D, false>::run(col);
#pragma unroll
        for (int al = 0; al < D; ++al) v[al][be] = col[al];
    }
}
template <int D>
TFC_HD void combine_inv2(c2 (&v)[D][D], int kyA, int kxA, int kyB, int kxB) {
    constexpr int P = 64 * D;
    const c2 wy = cis_neg2((float)kyA / (float)P, (float)kyB / (float)P);
    const c2 wx = cis_neg2((float)kxA / (float)P, (float)kxB / (float)P);
#pragma unroll
    for (int be = 0; be < D; ++be) {
        c2 col[D];
#pragma unroll
        for (int al = 0; al < D; ++al) col[al] = v[al][be];
        Dft<D, true>::run(col);
        c2 w = wy;
#pragma unroll
        for (int p = 1; p < D; ++p) {
            col[p] = cmulc2(col[p], w);
            if (p + 1 < D) w = cmul2(w, wy);
        }
#pragma unroll
        for (int p = 0; p < D; ++p) v[p][be] = col[p];
    }
#pragma unroll
    for (int p = 0; p < D; ++p) {
        Dft<D, true>::run(v[p]);
        c2 w = wx;
#pragma unroll
        for (int q = 1; q < D; ++q) {
            v[p][q] = cmulc2(v[p][q], w);
            if (q + 1 < D) w = cmul2(w, wx);
        }
    }
}

// position pairs {(ky',kx'), -(ky',kx')} of the 64 x 64 sub-grid: 64 x 32 slots walk the half-plane column
// POSITIONS in memory order (4 consecutive float4 = 64 contiguous bytes per group; slot j = 0 is the kx' = 0
// column, handled with kx' = 32 by the 66 items that pair rows instead)
// Out-of-line copy of the packed bin evaluation: combine_item calls it from 8 (D = 4) fully unrolled entry
// pairs; inlining it there made the kernel instruction-fetch bound (ncu: no_instruction 3.9 per issue).
#ifdef __CUDA_ARCH__
__device__ __noinline__
#else
inline
#endif
c2 bin_eval_pair_call(const Params& prm, bool mse, bool phase, c2 zk, c2 zm, float2& accA, float2& accP) {
    return bin_eval_pair(prm, mse, phase, zk, zm, accA, accP);
}

#ifdef __CUDA_ARCH__
__device__ __noinline__
#else
inline
#endif
float2 bin_eval_call(const Params& prm, float2 zk, float2 zm, float mult, float& accA, float& accP) {
    return bin_eval(prm, zk, zm, mult, accA, accP);
}

constexpr int kCombineItems = 64 * 32 + 2 * 33;

// One position pair of one tile: load the D^2 sub-spectra at both positions, combine, evaluate every
// full-size half-plane bin they contain, un-combine the spectral gradient, store back.
template <int D>
TFC_HD void combine_item(const Params& prm, float4* ws_tile, int item, float& accA, float& accP) {
    constexpr int P = 64 * D, NPP = D * D / 2, HD = D / 2;
    int kyA, kxA;
    if (item < 64 * 32) {
        const int j = item & 31;
        if (j == 0) return;
        kyA = freq_of_pos<64>(item >> 5);
        kxA = freq_of_pos<64>((j >> 2) * 8 + (j & 3));  // 1 .. 31
    } else {
        const int sp = item - 64 * 32;
        kxA = (sp / 33) * 32;
        kyA = sp % 33;
    }
    const int kyB = (64 - kyA) & 63, kxB = (64 - kxA) & 63;
    const bool self = (kyA == kyB) && (kxA == kxB);
    const int offA = pos_of_freq<64>(kyA) * 64 + pos_of_freq<64>(kxA);
    const int offB = pos_of_freq<64>(kyB) * 64 + pos_of_freq<64>(kxB);
    float2 za[D][D], zb[D][D];
#pragma unroll
    for (int pl = 0; pl < NPP; ++pl) {
        const int p = pl / HD, q = 2 * (pl % HD);
        const float4 a = ws_tile[pl * 4096 + offA];
        const float4 b = ws_tile[pl * 4096 + offB];
        za[p][q] = make_float2(a.x, a.z);
        za[p][q + 1] = make_float2(a.y, a.w);
        zb[p][q] = make_float2(b.x, b.z);
        zb[p][q + 1] = make_float2(b.y, b.w);
    }
    {
        c2 z2[D][D];
#pragma unroll
        for (int p = 0; p < D; ++p)
#pragma unroll
            for (int q = 0; q < D; ++q)
                z2[p][q] = make_c2(make_float2(za[p][q].x, zb[p][q].x), make_float2(za[p][q].y, zb[p][q].y));
        combine_fwd2<D>(z2, kyA, kxA, kyB, kxB);
#pragma unroll
        for (int p = 0; p < D; ++p)
#pragma unroll
            for (int q = 0; q < D; ++q) {
                za[p][q] = make_float2(z2[p][q].re.x, z2[p][q].im.x);
                zb[p][q] = make_float2(z2[p][q].re.y, z2[p][q].im.y);
            }
    }
    const bool want_grad = prm.grad != nullptr;
    const bool full = (prm.flags & TFCFFT_FULL_SPECTRUM) != 0;
    // The partner of the full frequency (kyA + 64 al, kxA + 64 be) sits in position B at (alB, beB) with
    // alB = kyA ? D-1-al : (D-al)%D (same for be).  Both maps are involutions; apply them once to zb so that the
    // entry loop below indexes registers statically.
    float2 zp[D][D];
#pragma unroll
    for (int al = 0; al < D; ++al)
#pragma unroll
        for (int be = 0; be < D; ++be) {
            const float2 r0 = kyA ? zb[D - 1 - al][be] : zb[(D - al) % D][be];
            zp[al][be] = r0;
        }
#pragma unroll
    for (int al = 0; al < D; ++al) {
        float2 t[D];
#pragma unroll
        for (int be = 0; be < D; ++be) t[be] = kxA ? zp[al][D - 1 - be] : zp[al][(D - be) % D];
#pragma unroll
        for (int be = 0; be < D; ++be) zp[al][be] = t[be];
    }
    float2 ga[D][D], gp[D][D];  // gp: gradient of the partner entries, in the permuted index space
    const bool generic = (prm.flags & (TFCFFT_LOG_MAGNITUDE | TFCFFT_FULL_SPECTRUM)) != 0;
    if (generic) {
#pragma unroll
        for (int al = 0; al < D; ++al)
#pragma unroll
            for (int be = 0; be < D; ++be) {
                const int alB = kyA ? D - 1 - al : (D - al) % D;
                const int beB = kxA ? D - 1 - be : (D - be) % D;
                float2 gk = make_float2(0.f, 0.f), gm = make_float2(0.f, 0.f);
                const bool skip = self && (al * D + be) > (alB * D + beB);  // each unordered pair once
                if (!skip) {
                    const bool selfbin = self && al == alB && be == beB;
                    const int kxf = kxA + 64 * be;
                    const float2 zk = za[al][be], zm = zp[al][be];
                    if (kxf == 0 || kxf == P / 2) {  // self-conjugate column: k and -k are both half-plane bins
                        gk = bin_eval_call(prm, zk, zm, 1.f, accA, accP);
                        if (!selfbin) gm = bin_eval_call(prm, zm, zk, 1.f, accA, accP);
                    } else if (kxf < P / 2) {
                        gk = bin_eval_call(prm, zk, zm, full ? 2.f : 1.f, accA, accP);
                    } else {
                        gm = bin_eval_call(prm, zm, zk, full ? 2.f : 1.f, accA, accP);
                    }
                }
                ga[al][be] = gk;
                gp[al][be] = gm;
            }
    } else {
        // default modes: two entries (be, be+1) per packed evaluation, MUFU / polynomial transcendental path
        const bool mse = (prm.flags & TFCFFT_DIST_MSE) != 0, phase = !(prm.flags & TFCFFT_NO_PHASE);
        float2 pA = make_float2(0.f, 0.f), pP = make_float2(0.f, 0.f);
        const float2 z0 = make_float2(0.f, 0.f);
#pragma unroll
        for (int al = 0; al < D; ++al)
#pragma unroll
            for (int bp = 0; bp < D / 2; ++bp) {
                const int alB = kyA ? D - 1 - al : (D - al) % D;
                bool isM[2], both[2], live[2];
                c2 zk, zm;
                float2 k_[2], m_[2];
#pragma unroll
                for (int l = 0; l < 2; ++l) {
                    const int be = 2 * bp + l;
                    const int beB = kxA ? D - 1 - be : (D - be) % D;
                    const int kxf = kxA + 64 * be;
                    live[l] = !(self && (al * D + be) > (alB * D + beB));
                    const bool special = (kxf == 0 || kxf == P / 2);
                    both[l] = live[l] && special && !(self && al == alB && be == beB);
                    isM[l] = !special && kxf > P / 2;
                    k_[l] = live[l] ? (isM[l] ? zp[al][be] : za[al][be]) : z0;
                    m_[l] = live[l] ? (isM[l] ? za[al][be] : zp[al][be]) : z0;
                }
                zk = make_c2(make_float2(k_[0].x, k_[1].x), make_float2(k_[0].y, k_[1].y));
                zm = make_c2(make_float2(m_[0].x, m_[1].x), make_float2(m_[0].y, m_[1].y));
                const c2 g = bin_eval_pair_call(prm, mse, phase, zk, zm, pA, pP);
                c2 g2 = make_c2(z0, z0);
                if (both[0] || both[1]) {  // self-conjugate columns only: evaluate the mirrored bin as well
                    const c2 zk2 = make_c2(make_float2(both[0] ? m_[0].x : 0.f, both[1] ? m_[1].x : 0.f),
                                           make_float2(both[0] ? m_[0].y : 0.f, both[1] ? m_[1].y : 0.f));
                    const c2 zm2 = make_c2(make_float2(both[0] ? k_[0].x : 0.f, both[1] ? k_[1].x : 0.f),
                                           make_float2(both[0] ? k_[0].y : 0.f, both[1] ? k_[1].y : 0.f));
                    g2 = bin_eval_pair_call(prm, mse, phase, zk2, zm2, pA, pP);
                }
#pragma unroll
                for (int l = 0; l < 2; ++l) {
                    const int be = 2 * bp + l;
                    const float2 gl = l ? make_float2(g.re.y, g.im.y) : make_float2(g.re.x, g.im.x);
                    const float2 g2l = l ? make_float2(g2.re.y, g2.im.y) : make_float2(g2.re.x, g2.im.x);
                    ga[al][be] = live[l] ? (isM[l] ? z0 : gl) : z0;
                    gp[al][be] = live[l] ? (isM[l] ? gl : (both[l] ? g2l : z0)) : z0;
                }
            }
        accA += pA.x + pA.y;
        accP += pP.x + pP.y;
    }
    if (!want_grad) return;
    // undo the permutation: gb[a][b] = gp[perm(a)][perm(b)]
    float2 gb[D][D];
#pragma unroll
    for (int al = 0; al < D; ++al) {
        float2 t[D];
#pragma unroll
        for (int be = 0; be < D; ++be) t[be] = kxA ? gp[al][D - 1 - be] : gp[al][(D - be) % D];
#pragma unroll
        for (int be = 0; be < D; ++be) gp[al][be] = t[be];
    }
#pragma unroll
    for (int al = 0; al < D; ++al)
#pragma unroll
        for (int be = 0; be < D; ++be) gb[al][be] = kyA ? gp[D - 1 - al][be] : gp[(D - al) % D][be];
    if (self) {
#pragma unroll
        for (int al = 0; al < D; ++al)
#pragma unroll
            for (int be = 0; be < D; ++be) ga[al][be] = cadd(ga[al][be], gb[al][be]);
    }
    {
        c2 g2[D][D];
#pragma unroll
        for (int p = 0; p < D; ++p)
#pragma unroll
            for (int q = 0; q < D; ++q)
                g2[p][q] = make_c2(make_float2(ga[p][q].x, gb[p][q].x), make_float2(ga[p][q].y, gb[p][q].y));
        combine_inv2<D>(g2, kyA, kxA, kyB, kxB);
#pragma unroll
        for (int p = 0; p < D; ++p)
#pragma unroll
            for (int q = 0; q < D; ++q) {
                ga[p][q] = make_float2(g2[p][q].re.x, g2[p][q].im.x);
                gb[p][q] = make_float2(g2[p][q].re.y, g2[p][q].im.y);
            }
    }
#pragma unroll
    for (int pl = 0; pl < NPP; ++pl) {
        const int p = pl / HD, q = 2 * (pl % HD);
        ws_tile[pl * 4096 + offA] = make_float4(ga[p][q].x, ga[p][q + 1].x, ga[p][q].y, ga[p][q + 1].y);
        if (!self) ws_tile[pl * 4096 + offB] = make_float4(gb[p][q].x, gb[p][q + 1].x, gb[p][q].y, gb[p][q + 1].y);
    }
}

TFC_HD bool sub_supported(const Params& prm) {
    return (prm.p == 128 || prm.p == 256) && prm.spec_mode == 0 && !(prm.flags & (TFCFFT_FORCE_SPLIT | TFCFFT_FORCE_GENERIC));
}

}  // namespace tfcfft
