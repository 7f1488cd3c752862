// combine8.cuh -- launch 2 of the sub-tile pipeline for 512 x 512 tiles (D = 8, 64 interleaved 64 x 64 sub-images).
//
// combine_kernel<2|4> (sub_tile.cuh) keeps all D^2 sub-spectra of a position pair in the registers of one thread;
// 64 sub-spectra x 2 positions do not fit.  The D x D butterfly is separable, so here one CTA stages the 64
// sub-spectra of up to 33 position pairs {k', -k'} in shared memory ([plane p*8+q][slot], 34 KB) and walks them in
// three phases of small per-thread jobs:
//   a  (pair, p):   twiddle W_P^{q kx'}, 8-point DFT over q            S[p][q]  -> T[p][be]
//   b  (pair, be):  twiddle W_P^{p ky'}, 8-point DFT over p            T[p][be] -> Z[al][be]; meets Z(k) with Z(-k)
//                   (column be of position A, column beB of position B), loss + spectral gradient with the packed
//                   bin evaluation, inverse 8-point DFT over al, conjugate twiddle             G[al][be] -> G'[p][be]
//   c  (pair, p):   inverse 8-point DFT over be, conjugate twiddle, Hermitian symmetrisation over the position
//                   pair and packing of two sub-image gradients per complex plane (as combine_item does)
// then writes the 32 packed planes back in place.  Both positions of a pair ride in the two lanes of the packed
// fp32 arithmetic (c2: lane x = position A, lane y = its partner B).  The arithmetic per entry is the same as
// combine_fwd2 / combine_item / combine_inv2 in the same order.
//
// CTA `row` (0..63) of a tile owns the pairs A = (row, kx'), B = (-row, -kx') for kx' = 1..31 and, when row <= 32,
// the two pairs of the self-conjugate columns kx' = 0 and 32 (A = (row, kx'), B = (-row, kx')).
#pragma once
#include "sub_tile.cuh"

namespace tfcfft {

struct Combine8Cfg {
    static constexpr int NT = 288;                 // >= 8 * 33 jobs: every phase is one round
    static constexpr int NPAIR = 33, NSLOT = 66;   // slots 0..32: positions A, 33..65: positions B
    static constexpr int PARTS = 64;               // CTAs (= partial sums) per tile
    static constexpr size_t SMEM = ((size_t)64 * NSLOT + 512) * sizeof(float2);  // staged sub-spectra + W_512 table
};

// pair j of CTA `row`: j = 0..30 -> kx' = j + 1; j = 31 -> kx' = 0; j = 32 -> kx' = 32 (rows <= 32 only)
struct Pair8 {
    int kyA, kxA, kyB, kxB, slotA, slotB;
    bool valid, self;
};
TFC_HD Pair8 pair8(int row, int j) {
    Pair8 r;
    r.kyA = row;
    r.kyB = (64 - row) & 63;
    if (j < 31) {
        r.kxA = j + 1;
        r.kxB = 63 - j;
        r.valid = true;
    } else {
        r.kxA = r.kxB = (j - 31) * 32;
        r.valid = row <= 32;
    }
    r.self = (r.kyA == r.kyB) && (r.kxA == r.kxB);
    r.slotA = j;
    r.slotB = r.self ? j : Combine8Cfg::NSLOT / 2 + j;
    return r;
}

// position `pos` (0..65) of CTA `row`: 0..32 = row `row`, kx' = pos; 33..65 = the partner row, kx' = 32..63, 0
struct Pos8 {
    int slot, off;
    bool ok;
};
TFC_HD Pos8 pos8(int row, int pos, int npair) {
    Pos8 r;
    const bool bside = pos >= 33;
    const int kx = bside ? ((pos - 1) & 63) : pos;
    const int j = (kx == 0) ? 31 : (kx == 32 ? 32 : (bside ? 63 - kx : kx - 1));
    const int kyB = (64 - row) & 63;
    const bool self = j >= 31 && kyB == row;
    r.ok = j < npair && !(bside && self);
    r.slot = bside ? Combine8Cfg::NSLOT / 2 + j : j;
    r.off = (bside ? kyB : row) * 64 + kx;
    return r;
}

// packed twiddle (lane x: W_512^ia, lane y: W_512^ib) from the shared-memory table
TFC_HD c2 tw2(const float2* tw, int ia, int ib) {
    const float2 a = tw[ia & 511], b = tw[ib & 511];
    return make_c2(make_float2(a.x, b.x), make_float2(a.y, b.y));
}

template <class Ctx>
TFC_HD void combine8_rows(const Ctx& ctx, const Params& prm, float2* ws_tile, int row, float2* sm, float& accA, float& accP,
                          const unsigned char* eqf = nullptr) {
    constexpr int D = 8, P = 512, NS = Combine8Cfg::NSLOT, NP = Combine8Cfg::NPAIR;
    const int npair = row <= 32 ? NP : NP - 2;
    const bool want_grad = prm.grad != nullptr;
    // the tile's 32 "fake == real" flag bytes (forward launch): loaded here, tested after the staging loads below
    unsigned long long fl[4] = {0ull, 0ull, 0ull, 0ull};
    if (eqf != nullptr) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
#ifdef __CUDA_ARCH__
            fl[i] = __ldcg(reinterpret_cast<const unsigned long long*>(eqf) + i);
#else
            for (int b = 0; b < 8; ++b) fl[i] |= (unsigned long long)eqf[8 * i + b] << (8 * b);
#endif
        }
    }
    // W_512^k table: the butterflies' twiddles W^{q kx'} / W^{p ky'} are looked up instead of being chained products
    float2* tw = sm + 64 * NS;
    for (int k = ctx.tid; k < P; k += ctx.nthreads) tw[k] = cis_neg((float)k / (float)P);
    // ---- load: 64 planes x (positions A, positions B).  Thread = (quarter of the planes, position): consecutive
    // threads read consecutive kx' of one plane row, 16 independent 8-byte loads in flight per thread ----
    for (int t = ctx.tid; t < 4 * NS; t += ctx.nthreads) {
        const int pg = t / NS;
        const Pos8 ps = pos8(row, t % NS, npair);
        if (ps.ok) {
            const float2* src = ws_tile + (pg * 16) * 4096 + ps.off;
            float2 v[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = ws_load(src + i * 4096);
#pragma unroll
            for (int i = 0; i < 16; ++i) sm[(pg * 16 + i) * NS + ps.slot] = v[i];
        }
    }
    ctx.sync();
    if ((fl[0] & fl[1] & fl[2] & fl[3]) == 0x0101010101010101ull) {
        // fake == real on the whole tile (uniform over the CTA): zero loss terms, zero-filled packed planes (see combine_item)
        if (want_grad) {
            for (int t = ctx.tid; t < 4 * NS; t += ctx.nthreads) {
                const int pg = t / NS;
                const Pos8 ps = pos8(row, t % NS, npair);
                if (ps.ok) {
                    float2* dst = ws_tile + (pg * 8) * 4096 + ps.off;
#pragma unroll
                    for (int i = 0; i < 8; ++i) dst[i * 4096] = make_float2(0.f, 0.f);
                }
            }
        }
        ctx.sync();
        return;
    }
    // ---- phase a: (p, pair): along q ----
    for (int it = ctx.tid; it < D * npair; it += ctx.nthreads) {
        const int p = it / npair, j = it % npair;
        const Pair8 pr = pair8(row, j);
        c2 v[D];
#pragma unroll
        for (int q = 0; q < D; ++q) {
            const float2 a = sm[(p * D + q) * NS + pr.slotA], b = sm[(p * D + q) * NS + pr.slotB];
            v[q] = make_c2(make_float2(a.x, b.x), make_float2(a.y, b.y));
        }
#pragma unroll
        for (int q = 1; q < D; ++q) v[q] = cmul2(v[q], tw2(tw, q * pr.kxA, q * pr.kxB));
        Dft<D, false>::run(v);
#pragma unroll
        for (int be = 0; be < D; ++be) {
            sm[(p * D + be) * NS + pr.slotA] = make_float2(v[be].re.x, v[be].im.x);
            if (!pr.self) sm[(p * D + be) * NS + pr.slotB] = make_float2(v[be].re.y, v[be].im.y);
        }
    }
    ctx.sync();
    // ---- phase b: (be, pair): along p, loss, spectral gradient, back along al ----
    const bool generic = (prm.flags & (TFCFFT_LOG_MAGNITUDE | TFCFFT_FULL_SPECTRUM)) != 0;
    const bool full = (prm.flags & TFCFFT_FULL_SPECTRUM) != 0;
    const bool mse = (prm.flags & TFCFFT_DIST_MSE) != 0, phase = !(prm.flags & TFCFFT_NO_PHASE);
    float2 pA = make_float2(0.f, 0.f), pP = make_float2(0.f, 0.f);
    const float2 z0 = make_float2(0.f, 0.f);
    for (int it = ctx.tid; it < D * npair; it += ctx.nthreads) {
        const int be = it / npair, j = it % npair;
        const Pair8 pr = pair8(row, j);
        const int beB = pr.kxA ? D - 1 - be : (D - be) % D;
        if (pr.self && be > beB) continue;  // the thread of the smaller column owns both columns of a self pair
        const bool same_col = pr.self && be == beB;
        const bool kyz = pr.kyA == 0;
        c2 col[D];
#pragma unroll
        for (int p = 0; p < D; ++p) {
            const float2 a = sm[(p * D + be) * NS + pr.slotA], b = sm[(p * D + beB) * NS + pr.slotB];
            col[p] = make_c2(make_float2(a.x, b.x), make_float2(a.y, b.y));
        }
#pragma unroll
        for (int p = 1; p < D; ++p) col[p] = cmul2(col[p], tw2(tw, p * pr.kyA, p * pr.kyB));
        Dft<D, false>::run(col);
        float2 za[D], zb[D], zp[D];
#pragma unroll
        for (int al = 0; al < D; ++al) {
            za[al] = make_float2(col[al].re.x, col[al].im.x);
            zb[al] = make_float2(col[al].re.y, col[al].im.y);
        }
        // partner of entry (al, be) of position A: entry (alB, beB) of position B, alB = ky' ? 7 - al : (8 - al) % 8
#pragma unroll
        for (int al = 0; al < D; ++al) zp[al] = kyz ? zb[(D - al) % D] : zb[D - 1 - al];
        const int kxf = pr.kxA + 64 * be;
        const bool special = (kxf == 0 || kxf == P / 2);  // self-conjugate column: k and -k are both half-plane bins
        const bool isM = !special && kxf > P / 2;         // the half-plane bin of the pair is the partner's
        float2 ga[D], gp[D];
        bool live[D], both[D];
        float2 k_[D], m_[D];
#pragma unroll
        for (int al = 0; al < D; ++al) {
            const int alB = kyz ? (D - al) % D : D - 1 - al;
            live[al] = !(same_col && al > alB);  // each unordered pair once
            both[al] = live[al] && special && !(same_col && al == alB);
            k_[al] = live[al] ? (isM ? zp[al] : za[al]) : z0;
            m_[al] = live[al] ? (isM ? za[al] : zp[al]) : z0;
        }
        if (generic) {
            float a = 0.f, p = 0.f;
#pragma unroll
            for (int al = 0; al < D; ++al) {
                float2 g = z0, g2 = z0;
                if (live[al]) {
                    g = bin_eval_call(prm, k_[al], m_[al], (full && !special) ? 2.f : 1.f, a, p);
                    if (both[al]) g2 = bin_eval_call(prm, m_[al], k_[al], 1.f, a, p);
                }
                ga[al] = isM ? z0 : g;
                gp[al] = isM ? g : g2;
            }
            pA.x += a;
            pP.x += p;
        } else {
            c2 g[D / 2], g2[D / 2];
#ifdef TFC_EVAL_OCT
            {
                c2 zk[4], zm[4];
#pragma unroll
                for (int h = 0; h < 4; ++h) {
                    const int al = 2 * h;
                    zk[h] = make_c2(make_float2(k_[al].x, k_[al + 1].x), make_float2(k_[al].y, k_[al + 1].y));
                    zm[h] = make_c2(make_float2(m_[al].x, m_[al + 1].x), make_float2(m_[al].y, m_[al + 1].y));
                }
                const OctEval q = bin_eval_oct_call(prm, mse, phase, zk[0], zm[0], zk[1], zm[1], zk[2], zm[2], zk[3], zm[3]);
#pragma unroll
                for (int h = 0; h < 4; ++h) g[h] = q.g[h];
                pA = p_add(pA, q.a);
                pP = p_add(pP, q.p);
            }
#else
#pragma unroll
            for (int e = 0; e < D / 2; e += 2) {
                c2 zk[2], zm[2];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int al = 2 * (e + h);
                    zk[h] = make_c2(make_float2(k_[al].x, k_[al + 1].x), make_float2(k_[al].y, k_[al + 1].y));
                    zm[h] = make_c2(make_float2(m_[al].x, m_[al + 1].x), make_float2(m_[al].y, m_[al + 1].y));
                }
                const QuadEval q = bin_eval_quad_call(prm, mse, phase, zk[0], zm[0], zk[1], zm[1]);
                g[e] = q.g0;
                g[e + 1] = q.g1;
                pA = p_add(pA, q.a);
                pP = p_add(pP, q.p);
            }
#endif
#pragma unroll
            for (int e = 0; e < D / 2; ++e) {
                const int al = 2 * e;
                g2[e] = make_c2(z0, z0);
                if (both[al] || both[al + 1]) {  // self-conjugate columns only: the mirrored bin as well
                    const c2 zk2 = make_c2(make_float2(both[al] ? m_[al].x : 0.f, both[al + 1] ? m_[al + 1].x : 0.f),
                                           make_float2(both[al] ? m_[al].y : 0.f, both[al + 1] ? m_[al + 1].y : 0.f));
                    const c2 zm2 = make_c2(make_float2(both[al] ? k_[al].x : 0.f, both[al + 1] ? k_[al + 1].x : 0.f),
                                           make_float2(both[al] ? k_[al].y : 0.f, both[al + 1] ? k_[al + 1].y : 0.f));
                    g2[e] = bin_eval_pair_call(prm, mse, phase, zk2, zm2, pA, pP);
                }
#pragma unroll
                for (int l = 0; l < 2; ++l) {
                    const float2 gl = l ? make_float2(g[e].re.y, g[e].im.y) : make_float2(g[e].re.x, g[e].im.x);
                    const float2 g2l = l ? make_float2(g2[e].re.y, g2[e].im.y) : make_float2(g2[e].re.x, g2[e].im.x);
                    ga[al + l] = live[al + l] ? (isM ? z0 : gl) : z0;
                    gp[al + l] = live[al + l] ? (isM ? gl : (both[al + l] ? g2l : z0)) : z0;
                }
            }
        }
        if (!want_grad) continue;
        // gradient w.r.t. the entries of position B's column: undo the partner permutation
        float2 gb[D];
#pragma unroll
        for (int al = 0; al < D; ++al) gb[al] = kyz ? gp[(D - al) % D] : gp[D - 1 - al];
        if (same_col) {
#pragma unroll
            for (int al = 0; al < D; ++al) ga[al] = cadd(ga[al], gb[al]);
        }
#pragma unroll
        for (int al = 0; al < D; ++al) col[al] = make_c2(make_float2(ga[al].x, gb[al].x), make_float2(ga[al].y, gb[al].y));
        Dft<D, true>::run(col);
#pragma unroll
        for (int p = 1; p < D; ++p) col[p] = cmulc2(col[p], tw2(tw, p * pr.kyA, p * pr.kyB));
#pragma unroll
        for (int p = 0; p < D; ++p) {
            sm[(p * D + be) * NS + pr.slotA] = make_float2(col[p].re.x, col[p].im.x);
            if (!same_col) sm[(p * D + beB) * NS + pr.slotB] = make_float2(col[p].re.y, col[p].im.y);
        }
    }
    accA += pA.x + pA.y;
    accP += pP.x + pP.y;
    if (!want_grad) return;
    ctx.sync();
    // ---- phase c: (p, pair): back along be, Hermitian symmetrisation, packing ----
    for (int it = ctx.tid; it < D * npair; it += ctx.nthreads) {
        const int p = it / npair, j = it % npair;
        const Pair8 pr = pair8(row, j);
        c2 v[D];
#pragma unroll
        for (int be = 0; be < D; ++be) {
            const float2 a = sm[(p * D + be) * NS + pr.slotA], b = sm[(p * D + be) * NS + pr.slotB];
            v[be] = make_c2(make_float2(a.x, b.x), make_float2(a.y, b.y));
        }
        Dft<D, true>::run(v);
#pragma unroll
        for (int q = 1; q < D; ++q) v[q] = cmulc2(v[q], tw2(tw, q * pr.kxA, q * pr.kxB));
#pragma unroll
        for (int i = 0; i < D / 2; ++i) {
            const float2 a0 = make_float2(v[2 * i].re.x, v[2 * i].im.x), b0 = make_float2(v[2 * i].re.y, v[2 * i].im.y);
            const float2 a1 = make_float2(v[2 * i + 1].re.x, v[2 * i + 1].im.x), b1 = make_float2(v[2 * i + 1].re.y, v[2 * i + 1].im.y);
            float2 h0, h1;  // Hs_{p,2i}(A), Hs_{p,2i+1}(A); Hs(B) = conj Hs(A)
            if (pr.self) {
                h0 = make_float2(a0.x, 0.f);
                h1 = make_float2(a1.x, 0.f);
            } else {
                h0 = make_float2(0.5f * (a0.x + b0.x), 0.5f * (a0.y - b0.y));
                h1 = make_float2(0.5f * (a1.x + b1.x), 0.5f * (a1.y - b1.y));
            }
            // packed plane p*4+i of the inverse launch, staged in rows 0..3 of this p (this thread owns row p of the pair)
            sm[(p * D + i) * NS + pr.slotA] = make_float2(h0.x - h1.y, h0.y + h1.x);
            if (!pr.self) sm[(p * D + i) * NS + pr.slotB] = make_float2(h0.x + h1.y, h1.x - h0.y);
        }
    }
    ctx.sync();
    // ---- store: 32 packed planes, same position walk as the load (8 planes per thread) ----
    for (int t = ctx.tid; t < 4 * NS; t += ctx.nthreads) {
        const int pg = t / NS;
        const Pos8 ps = pos8(row, t % NS, npair);
        if (ps.ok) {
            float2* dst = ws_tile + (pg * 8) * 4096 + ps.off;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int pl = pg * 8 + i;
                dst[i * 4096] = sm[((pl >> 2) * D + (pl & 3)) * NS + ps.slot];
            }
        }
    }
    ctx.sync();
}

}  // namespace tfcfft
