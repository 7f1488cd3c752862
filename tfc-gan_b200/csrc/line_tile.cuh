// line_tile.cuh -- 64 x 64 tiles with one thread per transform line (register-resident 64-point FFTs).
//
// Why: the packed pair kernel (pair_tile.cuh) is bounded by shared-memory round trips -- every radix-8 pass
// moves the whole tile through shared memory and ends in a block barrier (profiles/r01_*: ~11.5 k shared
// wavefronts and 9 barriers per pair).  Here a thread owns a whole 64-point line: it reads the line once,
// runs the complete 8 x 8 Cooley-Tukey factorisation in registers with compile-time twiddles (no twiddle
// table, no index arithmetic, natural-order output) and writes it back once.  Shared-memory traffic per tile
// drops ~2.3x and the instruction count ~1.8x; parallelism comes from instruction-level parallelism inside a
// thread (64 independent elements) and from six independent 64-thread CTAs per SM.
//
// One CTA (64 threads) = one tile: cooperative coalesced load + luma -> rows -> columns -> half-plane loss +
// spectral gradient -> inverse columns (33 non-zero ones) -> inverse rows -> coalesced gradient store.
#pragma once
#include "pair_tile.cuh"

namespace tfcfft {

struct LineCfg {
    static constexpr int NT = 64;
    static constexpr int LD = 65;  // float2 row pitch: lanes = rows and lanes = columns are both conflict-free
    static constexpr size_t SMEM = (size_t)64 * LD * sizeof(float2);
#ifndef TFCFFT_LINE_BIN_EVALS
#define TFCFFT_LINE_BIN_EVALS 1
#endif
    static constexpr int BIN_EVALS = TFCFFT_LINE_BIN_EVALS;  // packed bin evaluations in flight per thread (2 measured no faster)
};

constexpr float kCos64[64] = {
    1.f, 0.995184727f, 0.98078528f, 0.956940336f, 0.923879533f, 0.881921264f, 0.831469612f, 0.773010453f,
    0.707106781f, 0.634393284f, 0.555570233f, 0.471396737f, 0.382683432f, 0.290284677f, 0.195090322f, 0.0980171403f,
    0.f, -0.0980171403f, -0.195090322f, -0.290284677f, -0.382683432f, -0.471396737f, -0.555570233f, -0.634393284f,
    -0.707106781f, -0.773010453f, -0.831469612f, -0.881921264f, -0.923879533f, -0.956940336f, -0.98078528f, -0.995184727f,
    -1.f, -0.995184727f, -0.98078528f, -0.956940336f, -0.923879533f, -0.881921264f, -0.831469612f, -0.773010453f,
    -0.707106781f, -0.634393284f, -0.555570233f, -0.471396737f, -0.382683432f, -0.290284677f, -0.195090322f, -0.0980171403f,
    0.f, 0.0980171403f, 0.195090322f, 0.290284677f, 0.382683432f, 0.471396737f, 0.555570233f, 0.634393284f,
    0.707106781f, 0.773010453f, 0.831469612f, 0.881921264f, 0.923879533f, 0.956940336f, 0.98078528f, 0.995184727f};
constexpr float kSin64[64] = {
    0.f, 0.0980171403f, 0.195090322f, 0.290284677f, 0.382683432f, 0.471396737f, 0.555570233f, 0.634393284f,
    0.707106781f, 0.773010453f, 0.831469612f, 0.881921264f, 0.923879533f, 0.956940336f, 0.98078528f, 0.995184727f,
    1.f, 0.995184727f, 0.98078528f, 0.956940336f, 0.923879533f, 0.881921264f, 0.831469612f, 0.773010453f,
    0.707106781f, 0.634393284f, 0.555570233f, 0.471396737f, 0.382683432f, 0.290284677f, 0.195090322f, 0.0980171403f,
    0.f, -0.0980171403f, -0.195090322f, -0.290284677f, -0.382683432f, -0.471396737f, -0.555570233f, -0.634393284f,
    -0.707106781f, -0.773010453f, -0.831469612f, -0.881921264f, -0.923879533f, -0.956940336f, -0.98078528f, -0.995184727f,
    -1.f, -0.995184727f, -0.98078528f, -0.956940336f, -0.923879533f, -0.881921264f, -0.831469612f, -0.773010453f,
    -0.707106781f, -0.634393284f, -0.555570233f, -0.471396737f, -0.382683432f, -0.290284677f, -0.195090322f, -0.0980171403f};

// a * W_64^K (INV: conjugate twiddle), K compile-time
template <int K, bool INV>
TFC_HD float2 mul_w64(float2 a) {
    constexpr int k = K & 63;
    constexpr float c8 = 0.707106781186547524f;
    if constexpr (k == 0) return a;
    else if constexpr (k == 16) return INV ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x);
    else if constexpr (k == 32) return make_float2(-a.x, -a.y);
    else if constexpr (k == 48) return INV ? make_float2(a.y, -a.x) : make_float2(-a.y, a.x);
    else if constexpr (k == 8) return INV ? make_float2((a.x - a.y) * c8, (a.x + a.y) * c8) : make_float2((a.x + a.y) * c8, (a.y - a.x) * c8);
    else {
        constexpr float wr = kCos64[k];
        constexpr float wi = INV ? kSin64[k] : -kSin64[k];
        return make_float2(a.x * wr - a.y * wi, a.x * wi + a.y * wr);
    }
}

template <bool INV, int N1, int K2>
TFC_HD void fft64_twiddle_row(float2* v) {  // v[N1 + 8*k2] *= W64^{N1*k2} for k2 = K2 .. 7
    if constexpr (K2 < 8) {
        v[N1 + 8 * K2] = mul_w64<N1 * K2, INV>(v[N1 + 8 * K2]);
        fft64_twiddle_row<INV, N1, K2 + 1>(v);
    }
}
template <bool INV, int N1>
TFC_HD void fft64_twiddles(float2* v) {
    if constexpr (N1 < 8) {
        fft64_twiddle_row<INV, N1, 1>(v);
        fft64_twiddles<INV, N1 + 1>(v);
    }
}

// In-register 64-point DFT.  Input natural order v[n]; output X[k2 + 8*k1] is left in slot k1 + 8*k2
// (i.e. slot s holds frequency fft64_freq(s) = (s >> 3) + 8 * (s & 7)); callers rename when they store.
template <bool INV>
TFC_HD void fft64(float2* v) {
    // step 1: for each n1, 8-point DFT over n2 of x[n1 + 8 n2]  ->  Y[n1][k2] in slot n1 + 8 k2
#pragma unroll
    for (int n1 = 0; n1 < 8; ++n1) {
        float2 u[8];
#pragma unroll
        for (int n2 = 0; n2 < 8; ++n2) u[n2] = v[n1 + 8 * n2];
        Dft<8, INV>::run(u);
#pragma unroll
        for (int k2 = 0; k2 < 8; ++k2) v[n1 + 8 * k2] = u[k2];
    }
    // step 2: Y[n1][k2] *= W64^{n1 k2}
    fft64_twiddles<INV, 1>(v);
    // step 3: for each k2, 8-point DFT over n1  ->  X[k2 + 8 k1] in slot k1 + 8 k2
#pragma unroll
    for (int k2 = 0; k2 < 8; ++k2) Dft<8, INV>::run(v + 8 * k2);
}
TFC_HD constexpr int fft64_freq(int slot) { return (slot >> 3) + 8 * (slot & 7); }

// physical slot of pixel x inside a freshly loaded row: the loader's four-pixel 64-bit stores of a half-warp
// land on 16 consecutive slots
TFC_HD constexpr int line_slot(int x) { return (x >> 2) + 16 * (x & 3); }

// ---- load: global -> luma -> complex tile (z = fake + i real) ------------------------------------
template <typename T, bool LUMA3, class Ctx>
TFC_HD void line_load(const Ctx& ctx, const Params& prm, const TileCoord& tc, float2* s) {
    constexpr int LD = LineCfg::LD, NC = LUMA3 ? 3 : 1;
    const T* fp = tile_ptr<T>(prm.fake, prm.fs, tc, 64);
    const T* rp = real_tile_ptr<T>(prm, tc, 64);
    const int fsh = (int)prm.fs[2], fsc = (int)prm.fs[1], rsh = (int)prm.rs[2], rsc = (int)prm.rs[1];
    const bool quant = (prm.flags & TFCFFT_QUANTIZE_U8) != 0;
    constexpr int NI = 4;  // items in flight per thread: 2*NC*NI 128-bit loads (the FFT registers are idle here)
#pragma unroll 1
    for (int it0 = ctx.tid; it0 < 64 * 16; it0 += NI * ctx.nthreads) {
        float raw[NI][2][NC][4];  // [item][fake|real][channel][pixel]
#pragma unroll
        for (int u = 0; u < NI; ++u) {
            const int it = it0 + u * ctx.nthreads, x = (it & 15) * 4, y = it >> 4;
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                IO<T>::load4(fp + y * fsh + c * fsc + x, raw[u][0][c]);
                IO<T>::load4(rp + y * rsh + c * rsc + x, raw[u][1][c]);
            }
        }
#pragma unroll
        for (int u = 0; u < NI; ++u) {
            const int it = it0 + u * ctx.nthreads, x4 = it & 15, y = it >> 4;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float z[2];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    if (!quant) {
                        float f = prm.lw[0] * raw[u][h][0][i];
                        if constexpr (LUMA3) f = fmaf(prm.lw[2], raw[u][h][2][i], fmaf(prm.lw[1], raw[u][h][1][i], f));
                        z[h] = f;
                    } else if constexpr (LUMA3) {
                        z[h] = (float)((19595 * IO<T>::quant(raw[u][h][0][i]) + 38470 * IO<T>::quant(raw[u][h][1][i]) +
                                        7471 * IO<T>::quant(raw[u][h][2][i]) + 0x8000) >> 16);
                    } else {
                        z[h] = (float)IO<T>::quant(raw[u][h][0][i]);
                    }
                }
                s[y * LD + x4 + 16 * i] = make_float2(z[0], z[1]);  // == line_slot(4*x4 + i)
            }
        }
    }
}

// ---- forward rows / columns: one thread per line ----------------------------------------------------
template <class Ctx>
TFC_HD void line_rows_fwd(const Ctx& ctx, float2* s) {
    constexpr int LD = LineCfg::LD;
    for (int y = ctx.tid; y < 64; y += ctx.nthreads) {
        float2* row = s + y * LD;
        float2 v[64];
#pragma unroll
        for (int x = 0; x < 64; ++x) v[x] = row[line_slot(x)];
        fft64<false>(v);
#pragma unroll
        for (int sl = 0; sl < 64; ++sl) row[fft64_freq(sl)] = v[sl];
    }
}
template <class Ctx>
TFC_HD void line_cols_fwd(const Ctx& ctx, float2* s) {
    constexpr int LD = LineCfg::LD;
    for (int x = ctx.tid; x < 64; x += ctx.nthreads) {
        float2* col = s + x;
        float2 v[64];
#pragma unroll
        for (int y = 0; y < 64; ++y) v[y] = col[y * LD];
        fft64<false>(v);
#pragma unroll
        for (int sl = 0; sl < 64; ++sl) col[fft64_freq(sl) * LD] = v[sl];
    }
}

// ---- loss + spectral gradient over the half plane (natural frequency order) ------------------------
template <class Ctx>
TFC_HD void line_bins(const Ctx& ctx, const Params& prm, float2* s, float& accA, float& accP) {
    constexpr int LD = LineCfg::LD;
    const bool mse = (prm.flags & TFCFFT_DIST_MSE) != 0, phase = !(prm.flags & TFCFFT_NO_PHASE);
    const bool want_grad = prm.grad != nullptr;
    const float2 z0 = make_float2(0.f, 0.f);
    float2 pA = z0, pP = z0;
    // two regular bins per packed evaluation, NE independent packed evaluations per iteration (their MUFU /
    // polynomial dependency chains interleave).  Kept as a real loop: the kernel is instruction-fetch sensitive
    // (straight-line 64-point transforms), the loss code should not be replicated 16 times
    constexpr int NE = LineCfg::BIN_EVALS;
    // a thread owns spectrum row ky and walks the columns kx = 1..31 (Z(k) at rk[kx], Z(-k) at rm[-kx]): two pointers,
    // no per-bin index arithmetic; consecutive threads = consecutive rows (odd pitch: conflict-free both ways)
    // work item t = spectrum row ky (t & 63) x part of the column range: groups of 64 threads take all of kx = 1..31,
    // groups of 128 split it 1..16 | 17..31
    const int parts = ctx.nthreads >= 128 ? 2 : 1;
    for (int t = ctx.tid; t < 64 * parts; t += ctx.nthreads) {
        const int ky = t & 63, part = t >> 6;
        float2* rk = s + ky * LD;
        float2* rm = s + ((64 - ky) & 63) * LD + 64;
        const int kx_lo = 1 + 16 * part, kx_hi = parts == 2 ? (part ? 32 : 17) : 32;
#pragma unroll 1
        for (int kx0 = kx_lo; kx0 < kx_hi; kx0 += 2 * NE) {
            float2 zk[NE][2], zm[NE][2];
            bool live[NE][2];
#pragma unroll
            for (int e = 0; e < NE; ++e)
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int kx = kx0 + 2 * e + u;
                    live[e][u] = kx < kx_hi;
                    zk[e][u] = live[e][u] ? rk[kx] : z0;
                    zm[e][u] = live[e][u] ? rm[-kx] : z0;
                }
            c2 g[NE];
            float2 qA[NE], qP[NE];
#pragma unroll
            for (int e = 0; e < NE; ++e) {
                qA[e] = z0;
                qP[e] = z0;
                g[e] = bin_eval_pair(prm, mse, phase, make_c2(make_float2(zk[e][0].x, zk[e][1].x), make_float2(zk[e][0].y, zk[e][1].y)),
                                     make_c2(make_float2(zm[e][0].x, zm[e][1].x), make_float2(zm[e][0].y, zm[e][1].y)), qA[e], qP[e]);
            }
#pragma unroll
            for (int e = 0; e < NE; ++e) {
                pA = p_add(pA, qA[e]);
                pP = p_add(pP, qP[e]);
                if (want_grad) {
                    const int kx = kx0 + 2 * e;
                    if (live[e][0]) {
                        rk[kx] = make_float2(g[e].re.x, g[e].im.x);
                        rm[-kx] = z0;
                    }
                    if (live[e][1]) {
                        rk[kx + 1] = make_float2(g[e].re.y, g[e].im.y);
                        rm[-kx - 1] = z0;
                    }
                }
            }
        }
    }
    // self-conjugate columns kx = 0 and kx = 32: item r owns rows (ky, -ky) = (r, 64 - r) of BOTH columns (r == 0: the
    // two self-conjugate rows 0 and 32); the two bins of a column are the two packed lanes.
    // Only the REAL part of these two columns' inverse column transforms reaches the (real) gradient, so their
    // spectral gradients are Hermitian-symmetrised, Hs(ky) = (G(ky) + conj G(-ky)) / 2, and the two real-output
    // transforms are packed into ONE complex line c = Hs_0 + i Hs_32 stored in column 0: the inverse column phase
    // then has exactly 32 lines (one warp) and leaves (u_0(y), u_32(y)) in column 0.
    for (int r = ctx.tid; r < 32; r += ctx.nthreads) {
        const int ky = r, kym = r == 0 ? 32 : 64 - r;
        float2 hs[2][2];  // [column][row ky | row kym]
#pragma unroll
        for (int cidx = 0; cidx < 2; ++cidx) {
            const int kx = cidx * 32;
            const float2 a = s[ky * LD + kx], b = s[kym * LD + kx];
            const float2 pa = r == 0 ? a : b, pb = r == 0 ? b : a;
            const c2 g = bin_eval_pair(prm, mse, phase, make_c2(make_float2(a.x, b.x), make_float2(a.y, b.y)),
                                       make_c2(make_float2(pa.x, pb.x), make_float2(pa.y, pb.y)), pA, pP);
            const float2 gk = make_float2(g.re.x, g.im.x), gm = make_float2(g.re.y, g.im.y);
            if (r == 0) {  // rows 0 and 32 are their own mirrors: Hs = Re G
                hs[cidx][0] = make_float2(gk.x, 0.f);
                hs[cidx][1] = make_float2(gm.x, 0.f);
            } else {
                hs[cidx][0] = make_float2(0.5f * (gk.x + gm.x), 0.5f * (gk.y - gm.y));
                hs[cidx][1] = make_float2(hs[cidx][0].x, -hs[cidx][0].y);
            }
        }
        if (want_grad) {  // c = Hs_0 + i Hs_32
            s[ky * LD] = make_float2(hs[0][0].x - hs[1][0].y, hs[0][0].y + hs[1][0].x);
            s[kym * LD] = make_float2(hs[0][1].x - hs[1][1].y, hs[0][1].y + hs[1][1].x);
        }
    }
    accA += pA.x + pA.y;
    accP += pP.x + pP.y;
}

// ---- inverse columns (32 lines: kx = 1..31 and the packed pair {0, 32}) and rows -------------------------------------------------------
template <class Ctx>
TFC_HD void line_cols_inv(const Ctx& ctx, float2* s) {
    constexpr int LD = LineCfg::LD;
    for (int x = ctx.tid; x < 32; x += ctx.nthreads) {  // line 0 = the packed columns 0 and 32
        float2* col = s + x;
        float2 v[64];
#pragma unroll
        for (int y = 0; y < 64; ++y) v[y] = col[y * LD];
        fft64<true>(v);
#pragma unroll
        for (int sl = 0; sl < 64; ++sl) col[fft64_freq(sl) * LD] = v[sl];
    }
}
// rows: columns 33..63 hold exact zeros (written by line_bins); the real part is the gradient, left in the
// row as floats g[x] at float index x (the row is reused as a float array)
template <class Ctx>
TFC_HD void line_rows_inv(const Ctx& ctx, float2* s) {
    constexpr int LD = LineCfg::LD;
    for (int y = ctx.tid; y < 64; y += ctx.nthreads) {
        float2* row = s + y * LD;
        float2 v[64];
#pragma unroll
        for (int k = 0; k < 64; ++k) v[k] = (k >= 1 && k < 32) ? row[k] : make_float2(0.f, 0.f);
        {
            const float2 u = row[0];  // (u_0(y), u_32(y)): real inverse transforms of columns 0 and 32
            v[0] = make_float2(u.x, 0.f);
            v[32] = make_float2(u.y, 0.f);
        }
        fft64<true>(v);
        float* g = reinterpret_cast<float*>(row);
#pragma unroll
        for (int m = 0; m < 32; ++m) {  // pixel x lives in slot (x >> 3) + 8 * (x & 7)
            const int x0 = 2 * m, x1 = 2 * m + 1;
            const float a = v[(x0 >> 3) + 8 * (x0 & 7)].x, b = v[(x1 >> 3) + 8 * (x1 & 7)].x;
            *reinterpret_cast<float2*>(g + 2 * m) = make_float2(a, b);
        }
    }
}

// ---- all four transform passes on ONE copy of the 64-point core ---------------------------------------------------
// pass 0: forward rows, 1: forward columns, 2: inverse columns (32 lines: kx = 1..31 and the packed pair {0, 32}),
// 3: inverse rows (real part -> floats).  The unnormalised inverse is the forward transform read backwards,
//     IDFT(x)[n] = DFT(x)[(64 - n) mod 64],
// so every pass runs the SAME straight-line `fft64<false>` code and differs only in where it loads and stores.  The
// caller keeps the pass loop rolled (`#pragma unroll 1`): the kernel then holds one ~9 KB copy of the core instead of
// four specialised ones (~52 KB), which is what the instruction cache of an SM can keep hot while several tiles are in
// different passes (ncu, profiles/r02_ncu_ring_4x7_summary.txt: `no_instruction` was the top stall, 55 % of it inside
// the forward transforms).  Cost: the zero-pruned / real-output shortcuts of the specialised inverse rows are gone
// (+2.5 % executed instructions).
template <class Ctx>
TFC_HD void line_fft_pass(const Ctx& ctx, float2* s, int pass) {
    constexpr int LD = LineCfg::LD;
    const int nlines = pass == 2 ? 32 : 64;
    for (int l = ctx.tid; l < nlines; l += ctx.nthreads) {
        float2 v[64];
        float2* row = s + l * LD;
        float2* col = s + l;
        if (pass == 0) {
#pragma unroll
            for (int x = 0; x < 64; ++x) v[x] = row[line_slot(x)];
        } else if (pass == 3) {
            // columns 33..63 hold exact zeros and column 0 holds (u_0(y), u_32(y)), the real inverse transforms of
            // columns 0 and 32 (line_bins)
#pragma unroll
            for (int k = 0; k < 64; ++k) v[k] = (k >= 1 && k < 32) ? row[k] : make_float2(0.f, 0.f);
            const float2 u = row[0];
            v[0] = make_float2(u.x, 0.f);
            v[32] = make_float2(u.y, 0.f);
        } else {
#pragma unroll
            for (int y = 0; y < 64; ++y) v[y] = col[y * LD];
        }
        fft64<false>(v);
        if (pass == 0) {
#pragma unroll
            for (int sl = 0; sl < 64; ++sl) row[fft64_freq(sl)] = v[sl];
        } else if (pass == 1) {
#pragma unroll
            for (int sl = 0; sl < 64; ++sl) col[fft64_freq(sl) * LD] = v[sl];
        } else if (pass == 2) {
#pragma unroll
            for (int sl = 0; sl < 64; ++sl) col[((64 - fft64_freq(sl)) & 63) * LD] = v[sl];
        } else {
            // gradient row as floats g[x] at float index x (the row is reused as a float array): pixel x is the
            // real part of frequency (64 - x) mod 64, which lives in slot (f >> 3) + 8 * (f & 7)
            float* g = reinterpret_cast<float*>(row);
#pragma unroll
            for (int m = 0; m < 32; ++m) {
                const int f0 = (64 - 2 * m) & 63, f1 = (64 - (2 * m + 1)) & 63;
                const float a = v[(f0 >> 3) + 8 * (f0 & 7)].x, b = v[(f1 >> 3) + 8 * (f1 & 7)].x;
                *reinterpret_cast<float2*>(g + 2 * m) = make_float2(a, b);
            }
        }
    }
}

// ================================================================================================================
// Half-line engine: TWO threads per transform line (128 threads per tile).
//
// The thread-per-line engine holds 64 complex values per thread (~160 registers): a tile gets two warps and an SM
// whose shared memory holds four tiles runs eight compute warps -- ncu shows the fixed-latency `wait` stall on top and
// issue slots 43 % busy (profiles/r02_ncu_patch16_line_ring_final_summary.txt).  Here a line is split with one
// decimation-in-frequency radix-2 step across two threads (in different warps, no shuffles):
//     h = 0:  a[n] = x[n] + x[n+32]                 ->  X[2m]   = DFT32(a)[m]
//     h = 1:  b[n] = (x[n] - x[n+32]) * W64^n       ->  X[2m+1] = DFT32(b)[m]
// Each thread reads the whole line (smem bandwidth has head room: 23 % busy), keeps 32 complex values (~100
// registers) and writes its 32 outputs, so a tile gets FOUR warps at the same shared memory: twice the resident
// warps per SM to hide the fixed latencies.  All passes share one rolled copy of the 32-point core (`fft32<false>`;
// inverse passes store at mirrored indices, as in line_fft_pass); the half-zero input of the inverse rows halves its
// loads for free (x[n+32] = 0 for n >= 1).
// ================================================================================================================
template <bool INV, int N1, int K2>
TFC_HD void fft32_twiddle_row(float2* v) {  // v[N1 + 4*k2] *= W32^{N1*k2} for k2 = K2 .. 7
    if constexpr (K2 < 8) {
        v[N1 + 4 * K2] = mul_w<32, N1 * K2, INV>(v[N1 + 4 * K2]);
        fft32_twiddle_row<INV, N1, K2 + 1>(v);
    }
}
// In-register 32-point DFT = 4 x 8.  Input natural order v[n]; output X[k2 + 8*k1] is left in slot k1 + 4*k2.
template <bool INV>
TFC_HD void fft32(float2* v) {
#pragma unroll
    for (int n1 = 0; n1 < 4; ++n1) {
        float2 u[8];
#pragma unroll
        for (int n2 = 0; n2 < 8; ++n2) u[n2] = v[n1 + 4 * n2];
        Dft<8, INV>::run(u);
#pragma unroll
        for (int k2 = 0; k2 < 8; ++k2) v[n1 + 4 * k2] = u[k2];
    }
    fft32_twiddle_row<INV, 1, 1>(v);
    fft32_twiddle_row<INV, 2, 1>(v);
    fft32_twiddle_row<INV, 3, 1>(v);
#pragma unroll
    for (int k2 = 0; k2 < 8; ++k2) Dft<4, INV>::run(v + 4 * k2);
}
TFC_HD constexpr int fft32_freq(int slot) { return (slot >> 2) + 8 * (slot & 3); }

template <int N>
TFC_HD float2 half_in(float2 xl, float2 xh, int h) {  // radix-2 DIF input of half h
    return h ? mul_w64<N, false>(csub(xl, xh)) : cadd(xl, xh);
}
template <int N, int END, class Load>
TFC_HD void half_load(float2* a, int h, const Load& ld) {
    if constexpr (N < END) {
        a[N] = half_in<N>(ld(N), ld(N + 32), h);
        half_load<N + 1, END>(a, h, ld);
    }
}
template <int N, int END, class Load>
TFC_HD void half_load_lo(float2* a, int h, const Load& ld) {  // x[n + 32] == 0
    if constexpr (N < END) {
        a[N] = h ? mul_w64<N, false>(ld(N)) : ld(N);
        half_load_lo<N + 1, END>(a, h, ld);
    }
}

// One half-line: load + radix-2 step + 32-point core.  a[sl] = X[2 * fft32_freq(sl) + h] of the line's 64-point DFT.
TFC_HD void half_line_compute(float2* s, int pass, int l, int h, float2* a) {
    constexpr int LD = LineCfg::LD;
    const float2* row = s + l * LD;
    const float2* col = s + l;
    if (pass == 0) {
        half_load<0, 32>(a, h, [&](int n) { return row[line_slot(n)]; });
    } else if (pass == 3) {
        // row[0] = (u_0(y), u_32(y)): real inverse transforms of columns 0 and 32; columns 33..63 are exact zeros
        const float2 u = row[0];
        a[0] = h ? make_float2(u.x - u.y, 0.f) : make_float2(u.x + u.y, 0.f);
        half_load_lo<1, 32>(a, h, [&](int n) { return row[n]; });
    } else {
        half_load<0, 32>(a, h, [&](int n) { return col[n * LD]; });
    }
    fft32<false>(a);
}
// position of frequency f = 2 F + h of a MIRRORED store (inverse passes): (64 - f) mod 64
TFC_HD int half_mirror(int F, int h) { return F == 0 ? (h ? 63 : 0) : 64 - 2 * F - h; }
TFC_HD void half_line_store(float2* s, int pass, int l, int h, const float2* a) {
    constexpr int LD = LineCfg::LD;
    float2* row = s + l * LD;
    float2* col = s + l;
    if (pass == 0) {
#pragma unroll
        for (int sl = 0; sl < 32; ++sl) row[2 * fft32_freq(sl) + h] = a[sl];
    } else if (pass == 1) {
#pragma unroll
        for (int sl = 0; sl < 32; ++sl) col[(2 * fft32_freq(sl) + h) * LD] = a[sl];
    } else if (pass == 2) {
#pragma unroll
        for (int sl = 0; sl < 32; ++sl) col[half_mirror(fft32_freq(sl), h) * LD] = a[sl];
    } else {
        float* g = reinterpret_cast<float*>(row);  // gradient row as floats, pixel x at float index x
#pragma unroll
        for (int sl = 0; sl < 32; ++sl) g[half_mirror(fft32_freq(sl), h)] = a[sl].x;
    }
}
// All four passes, 128 threads per tile (thread = line + 64 * half); ends with the data stored but NOT yet
// synchronised (the caller's barrier follows, as for line_fft_pass).  Both halves of a line read the whole line and
// write into it: a barrier separates the loads from the stores.
template <class Ctx>
TFC_HD void line2_fft_pass(const Ctx& ctx, float2* s, int pass) {
    const int nlines = pass == 2 ? 32 : 64;
    if (ctx.nthreads == 128) {
        const int l = ctx.tid & 63, h = ctx.tid >> 6;
        float2 a[32];
        if (l < nlines) half_line_compute(s, pass, l, h, a);
        ctx.sync();
        if (l < nlines) half_line_store(s, pass, l, h, a);
    } else {  // serial emulation: a line at a time, both halves before either store
        for (int l = ctx.tid; l < nlines; l += ctx.nthreads) {
            float2 a0[32], a1[32];
            half_line_compute(s, pass, l, 0, a0);
            half_line_compute(s, pass, l, 1, a1);
            half_line_store(s, pass, l, 0, a0);
            half_line_store(s, pass, l, 1, a1);
        }
    }
}

// ---- gradient store: rows of floats -> global, 128-bit stores ----------------------------------------
template <typename T, bool LUMA3, bool ACC, class Ctx>
TFC_HD void line_store_rows(const Ctx& ctx, const GradOut& go, T* gp, int sh, int sc, const float2* s) {
    constexpr int LD = LineCfg::LD, NC = LUMA3 ? 3 : 1;
    if (ctx.nthreads == 64 || ctx.nthreads == 128) {
        // a thread keeps its 4-pixel column and walks down the rows: one pointer bump per row, no index arithmetic
        const int x = (ctx.tid & 15) * 4, y0 = ctx.tid >> 4;
        const int rstep = ctx.nthreads / 16;  // rows per sweep of the group: 4 or 8
        T* p = gp + y0 * sh + x;
        const float* g = reinterpret_cast<const float*>(s + y0 * LD) + x;
#pragma unroll 4
        for (int j = 0; j < 64 / rstep; ++j) {
            const float2 lo = *reinterpret_cast<const float2*>(g), hi = *reinterpret_cast<const float2*>(g + 2);
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                float v[4] = {go.w[c] * lo.x, go.w[c] * lo.y, go.w[c] * hi.x, go.w[c] * hi.y};
                if constexpr (ACC) {
                    float o[4];
                    IO<T>::load4(p + c * sc, o);
#pragma unroll
                    for (int i = 0; i < 4; ++i) v[i] += o[i];
                }
                IO<T>::store4(p + c * sc, v);
            }
            p += rstep * sh;
            g += rstep * LD * 2;
        }
    } else {
        for (int it = ctx.tid; it < 64 * 16; it += ctx.nthreads) {
            const int x = (it & 15) * 4, y = it >> 4;
            const float* g = reinterpret_cast<const float*>(s + y * LD) + x;
            const float2 lo = *reinterpret_cast<const float2*>(g), hi = *reinterpret_cast<const float2*>(g + 2);
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                float v[4] = {go.w[c] * lo.x, go.w[c] * lo.y, go.w[c] * hi.x, go.w[c] * hi.y};
                GradOut g1 = go;
                g1.acc = ACC;
                grad_store4<T>(g1, gp + y * sh + c * sc + x, v);
            }
        }
    }
}
template <typename T, bool LUMA3, class Ctx>
TFC_HD void line_store(const Ctx& ctx, const Params& prm, const TileCoord& tc, const float2* s) {
    T* gp = const_cast<T*>(tile_ptr<T>(prm.grad, prm.gs, tc, 64));
    const int sh = (int)prm.gs[2], sc = (int)prm.gs[1];
    const GradOut go = grad_out(prm);
    if (go.acc) line_store_rows<T, LUMA3, true>(ctx, go, gp, sh, sc, s);
    else line_store_rows<T, LUMA3, false>(ctx, go, gp, sh, sc, s);
}

// Pull the next tile's source lines into L2 shortly before they are needed (late enough to survive in L2, early
// enough to turn the load phase's four dependent HBM round trips into L2 hits).
template <typename T, bool LUMA3, class Ctx>
TFC_HD void line_prefetch_l2(const Ctx& ctx, const Params& prm, const TileCoord& tc) {
#ifdef __CUDA_ARCH__
    constexpr int NC = LUMA3 ? 3 : 1, EPL = 128 / (int)sizeof(T), LPR = (64 + EPL - 1) / EPL;
    const T* fp = tile_ptr<T>(prm.fake, prm.fs, tc, 64);
    const T* rp = real_tile_ptr<T>(prm, tc, 64);
    for (int it = ctx.tid; it < 2 * NC * 64 * LPR; it += ctx.nthreads) {
        const int l = it % LPR, y = (it / LPR) & 63, hc = it / (LPR * 64);
        const int h = hc & 1, c = hc >> 1;
        const T* q = (h ? rp + y * (int)prm.rs[2] + c * (int)prm.rs[1] : fp + y * (int)prm.fs[2] + c * (int)prm.fs[1]) + l * EPL;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(q));
    }
#endif
}

// ---- one tile ---------------------------------------------------------------------------------------
template <typename T, bool LUMA3, class Ctx>
TFC_HD void line_process(const Ctx& ctx, const Params& prm, int tile, float2* s, float& accA, float& accP,
                         int next_tile = -1, bool halfline = false) {
    const TileCoord tc = decode_tile(prm, tile);
    ctx.mark(0);
    line_load<T, LUMA3>(ctx, prm, tc, s);
    ctx.sync();
    ctx.mark(1);
    const bool want_grad = prm.grad != nullptr;
    const int npass = want_grad ? 4 : 2;
#pragma unroll 1
    for (int pass = 0; pass < npass; ++pass) {  // rolled: ONE copy of the 64-point core in the kernel
        if (halfline) line2_fft_pass(ctx, s, pass);  // CPU emulation of the half-line engine (TFCFFT_USE_HALFLINE)
        else line_fft_pass(ctx, s, pass);
        if (pass == 2 && next_tile >= 0) line_prefetch_l2<T, LUMA3>(ctx, prm, decode_tile(prm, next_tile));
        ctx.sync();
        ctx.mark(pass < 2 ? 2 + pass : 3 + pass);
        if (pass == 1) {
            line_bins(ctx, prm, s, accA, accP);
            ctx.sync();
            ctx.mark(4);
        }
    }
    if (want_grad) {
        ctx.mark(6);
        line_store<T, LUMA3>(ctx, prm, tc, s);
        ctx.sync();
        ctx.mark(7);
    }
}

}  // namespace tfcfft
