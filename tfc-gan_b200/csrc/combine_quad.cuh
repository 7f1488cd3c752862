// combine_quad.cuh -- launch 2 of the sub-tile pipeline for 256 x 256 tiles (D = 4), one position pair per thread QUAD.
//
// combine_item<4> (sub_tile.cuh) keeps the 16 sub-spectra of both positions of a pair {k', -k'} in ONE thread: ~2,500
// instructions in a row per item at 128 registers (spills around the evaluation calls), two items per thread, four
// warps per scheduler -- the launch is bounded by the latency of one thread's instruction stream (24 us at batch 64,
// 45 % issue utilisation; profiles/r02_final_ab.txt).  The 4 x 4 butterfly is separable, so here the four lanes of a
// quad share one item: lane r owns row p = r of the sub-spectra in the first stage and column be = r of the full-size
// spectrum entries afterwards.  Same arithmetic in the same order as combine_fwd2 / combine_item / combine_inv2:
//   load S[r][q] (q = 0..3) at both positions           twiddle W^{q kx'}, 4-point DFT over q        -> T[r][be]
//   transpose inside the quad (shared memory, warp-synchronous)                                      -> T[p][r]
//   twiddle W^{p ky'}, 4-point DFT over p                                                            -> Z[al][r]
//   the partner of entry (al, be) of position A sits in position B at (alB, 3 - be): one lane exchange with quad
//   lane 3 - r; lanes 0, 1 hold half-plane bins (kx < 128), lanes 2, 3 their mirrors (roles of k and -k swapped)
//   4 bin evaluations per lane (two packed calls of bin_eval_pair), spectral gradient back through the same exchange
//   inverse 4-point DFT over al, conjugate twiddle, transpose, inverse DFT over be, conjugate twiddle
//   Hermitian symmetrisation over the position pair, two sub-image gradients packed per complex plane (planes 2r, 2r+1)
// A thread runs ~650 instructions, 76 registers and no spills on its own.  MEASURED (profiles/r02_final_ab.txt): parity,
// not a win -- the same 10.7 M warp instructions per launch (every lane of a quad recomputes the row twiddle, the
// exchanges cost what the spills did), issue utilisation 49 % against 47 %, 28.8 against 27.3 us under ncu; the one-thread
// items of the self-conjugate columns, inlined into the same kernel, hold it at 96 registers = 5 CTAs per SM.  Opt-in
// (TFCFFT_COMBINE_QUAD=1); what would make it pay is a quad form of the self-conjugate columns (64 registers, 8 CTAs per SM).
//
// A CTA walks RPC rows of the position grid of one tile: in row `ky'` it owns the pairs A = (ky', kx'), B = (-ky', -kx') for
// kx' = 1..31 (quad kx' = 0 idles along); the 66 pairs of the self-conjugate columns kx' = 0 and 32 keep the one-thread
// code (combine_item<4>) and run in one extra CTA per tile (CombineQCfg::PARTS partial sums per tile).
#pragma once
#include "sub_tile.cuh"

namespace tfcfft {

struct CombineQCfg {
    static constexpr int NT = 128;                 // 4 warps x 8 quads = kx' 0..31 of one row of the position grid
    static constexpr int RPC = kCombineQRows;      // position rows per CTA, one after the other
    static constexpr int GROUPS = 64 / RPC;
    static constexpr int PARTS = kCombineQParts;   // CTAs (= partial sums) per tile: GROUPS + the self-conjugate columns
    static constexpr int XSTRIDE = 5;              // float4 per lane in the exchange area (4 + 1 pad: conflict-free)
    static constexpr int XWARP = 32 * XSTRIDE;     // float4 per warp
};

#ifdef __CUDACC__
// v[i] of quad lane r  ->  v[i] := (v[r] of quad lane i): a 4 x 4 transpose of packed complex pairs inside every quad
__device__ __forceinline__ void quad_transpose(c2 (&v)[4], float4* xw, int lane) {
    const int r = lane & 3, qb = lane & ~3;
#pragma unroll
    for (int i = 0; i < 4; ++i) xw[lane * CombineQCfg::XSTRIDE + i] = make_float4(v[i].re.x, v[i].re.y, v[i].im.x, v[i].im.y);
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float4 t = xw[(qb + i) * CombineQCfg::XSTRIDE + r];
        v[i] = make_c2(make_float2(t.x, t.y), make_float2(t.z, t.w));
    }
    __syncwarp();
}
__device__ __forceinline__ float2 shfl_xor2(float2 a, int m) {
    return make_float2(__shfl_xor_sync(0xffffffffu, a.x, m), __shfl_xor_sync(0xffffffffu, a.y, m));
}

// One position pair per quad; all 32 lanes of the warp call this together.  kxA in 1..31; an inactive quad (the kx' = 0
// slot of the row) computes along on a valid position but neither stores nor contributes to the loss sums.
// packed twiddle of a position pair: lane x = W^k (position A), lane y = W^{64 - k} = -i conj(W^k) (position B), W = e^{-2 pi i / 256}
__device__ __forceinline__ c2 quad_twiddle(int k) {
    const float2 a = cis_neg((float)k * (1.f / 256.f));
    return make_c2(make_float2(a.x, -a.y), make_float2(a.y, -a.x));
}

// this lane's sub-spectra S[r][q], q = 0..3, of the position pair (kyA, kxA) / its partner: a[q] = position A, b[q] = B
__device__ __forceinline__ void combine_quad_load(const float2* ws_tile, int kyA, int kxA, float2 (&a)[4], float2 (&b)[4]) {
    const int r = (int)threadIdx.x & 3;
    const int offA = kyA * 64 + kxA, offB = ((64 - kyA) & 63) * 64 + 64 - kxA;
    const float2* src = ws_tile + (r * 4) * 4096;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        a[q] = ws_load(src + q * 4096 + offA);
        b[q] = ws_load(src + q * 4096 + offB);
    }
}

// `wx` = quad_twiddle(kxA) (the same for every row a thread walks); a / b from combine_quad_load (the caller loads the next
// row's while this one is processed); `flags`: the tile's "fake == real" bytes of the forward launch
__device__ __forceinline__ void combine_quad_item(const Params& prm, float2* ws_tile, int kyA, int kxA, bool active, c2 wx,
                                                  unsigned long long flags, const float2 (&a)[4], const float2 (&b)[4], float4* xw,
                                                  float& accA, float& accP) {
    const int lane = (int)threadIdx.x & 31, r = lane & 3;
    const int kyB = (64 - kyA) & 63, kxB = 64 - kxA;
    const int offA = kyA * 64 + kxA, offB = kyB * 64 + kxB;
    c2 v[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) v[q] = make_c2(make_float2(a[q].x, b[q].x), make_float2(a[q].y, b[q].y));
    const bool want_grad = prm.grad != nullptr;
    if (flags == 0x0101010101010101ull) {  // identical tile (uniform over the CTA): zero terms, zero-filled planes
        if (want_grad && active) {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                float2* plane = ws_tile + (r * 2 + i) * 4096;
                plane[offA] = make_float2(0.f, 0.f);
                plane[offB] = make_float2(0.f, 0.f);
            }
        }
        return;
    }
    // ky' = 0 pairs with itself: W^0 on both lanes
    const c2 wy = kyA ? quad_twiddle(kyA) : make_c2(make_float2(1.f, 1.f), make_float2(0.f, 0.f));
    // ---- forward butterflies ----
    {
        c2 w = wx;
#pragma unroll
        for (int q = 1; q < 4; ++q) {
            v[q] = cmul2(v[q], w);
            if (q + 1 < 4) w = cmul2(w, wx);
        }
    }
    Dft<4, false>::run(v);          // T[r][be]
    quad_transpose(v, xw, lane);    // T[p][r]
    {
        c2 w = wy;
#pragma unroll
        for (int p = 1; p < 4; ++p) {
            v[p] = cmul2(v[p], w);
            if (p + 1 < 4) w = cmul2(w, wy);
        }
    }
    Dft<4, false>::run(v);          // Z[al][be = r]: lane x = position A, lane y = position B
    // ---- partners: entry (al, be) of A meets entry (alB, 3 - be) of B, alB = ky' ? 3 - al : (4 - al) % 4 ----
    float2 za[4], zp[4];
    {
        float2 zb[4];
#pragma unroll
        for (int al = 0; al < 4; ++al) {
            za[al] = make_float2(v[al].re.x, v[al].im.x);
            zb[al] = shfl_xor2(make_float2(v[al].re.y, v[al].im.y), 3);  // Z_B[al][3 - r]
        }
        if (kyA) {
#pragma unroll
            for (int al = 0; al < 4; ++al) zp[al] = zb[3 - al];
        } else {
#pragma unroll
            for (int al = 0; al < 4; ++al) zp[al] = zb[(4 - al) & 3];
        }
    }
    // ---- loss terms + spectral gradient: kx = kx' + 64 r is a half-plane bin for r < 2, the mirror of one otherwise ----
    const bool mir = r >= 2;
    const bool mse = (prm.flags & TFCFFT_DIST_MSE) != 0, phase = !(prm.flags & TFCFFT_NO_PHASE);
    float2 g[4];
    {
        float2 k_[4], m_[4];
#pragma unroll
        for (int al = 0; al < 4; ++al) {
            k_[al] = mir ? zp[al] : za[al];
            m_[al] = mir ? za[al] : zp[al];
        }
        float2 pA = make_float2(0.f, 0.f), pP = make_float2(0.f, 0.f);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const c2 zk = make_c2(make_float2(k_[2 * h].x, k_[2 * h + 1].x), make_float2(k_[2 * h].y, k_[2 * h + 1].y));
            const c2 zm = make_c2(make_float2(m_[2 * h].x, m_[2 * h + 1].x), make_float2(m_[2 * h].y, m_[2 * h + 1].y));
            const c2 gg = bin_eval_pair(prm, mse, phase, zk, zm, pA, pP);
            g[2 * h] = make_float2(gg.re.x, gg.im.x);
            g[2 * h + 1] = make_float2(gg.re.y, gg.im.y);
        }
        if (active) {
            accA += pA.x + pA.y;
            accP += pP.x + pP.y;
        }
    }
    if (!want_grad) return;
    // gradient w.r.t. this lane's A entries (ga) and, sent back through the same exchange, its B entries (gb)
    c2 col[4];
    {
        const float2 z0 = make_float2(0.f, 0.f);
        float2 rc[4];
#pragma unroll
        for (int al = 0; al < 4; ++al) rc[al] = shfl_xor2(mir ? g[al] : z0, 3);  // d / d Z_B[alB(al)][r], from lane 3 - r
#pragma unroll
        for (int al = 0; al < 4; ++al) {
            const float2 ga = mir ? z0 : g[al];
            const float2 gb = kyA ? rc[3 - al] : rc[(4 - al) & 3];
            col[al] = make_c2(make_float2(ga.x, gb.x), make_float2(ga.y, gb.y));
        }
    }
    // ---- inverse butterflies ----
    Dft<4, true>::run(col);
    {
        c2 w = wy;
#pragma unroll
        for (int p = 1; p < 4; ++p) {
            col[p] = cmulc2(col[p], w);
            if (p + 1 < 4) w = cmul2(w, wy);
        }
    }
    quad_transpose(col, xw, lane);  // H'[p = r][be]
    Dft<4, true>::run(col);
    {
        c2 w = wx;
#pragma unroll
        for (int q = 1; q < 4; ++q) {
            col[q] = cmulc2(col[q], w);
            if (q + 1 < 4) w = cmul2(w, wx);
        }
    }
    // ---- Hermitian symmetrisation over the pair, two sub-image gradients per complex plane (as combine_item) ----
    if (active) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const c2 e0 = col[2 * i], e1 = col[2 * i + 1];
            const float2 h0 = make_float2(0.5f * (e0.re.x + e0.re.y), 0.5f * (e0.im.x - e0.im.y));
            const float2 h1 = make_float2(0.5f * (e1.re.x + e1.re.y), 0.5f * (e1.im.x - e1.im.y));
            float2* plane = ws_tile + (r * 2 + i) * 4096;
            plane[offA] = make_float2(h0.x - h1.y, h0.y + h1.x);
            plane[offB] = make_float2(h0.x + h1.y, h1.x - h0.y);
        }
    }
}
#endif  // __CUDACC__

}  // namespace tfcfft
