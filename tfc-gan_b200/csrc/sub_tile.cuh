// sub_tile.cuh -- 128 x 128, 256 x 256 and 512 x 512 tiles on top of the thread-per-line 64 x 64 machinery.
//
// A P x P tile (P = 64 D, D = 2, 4 or 8; the D = 8 cross-sub-image step is combine8.cuh) is decimated in both dimensions into D x D interleaved sub-images
// s_pq[a, b] = x[D a + p, D b + q].  The 2-D DFT factors exactly (decimation in time):
//     Z[ky' + 64 al, kx' + 64 be] = sum_{p,q} W_D^{p al + q be} * W_P^{p ky' + q kx'} * S_pq[ky', kx']
// so the heavy work -- D^2 independent 64 x 64 complex transforms -- runs on the register-resident 64-point
// line transforms of line_tile.cuh, and the cross-sub-image step is a D x D butterfly per frequency position
// done in registers by `combine_kernel`, which also meets Z(k) with Z(-k), evaluates the loss and the spectral
// gradient, and applies the inverse butterfly.  Three launches exchange the sub-spectra through an L2-sized
// workspace chunk (P*P*8 bytes per tile, D^2 planes of 64 x 64 float2 in natural frequency order):
//   sub_fwd_kernel:      sub-images -> S_pq                      (HBM read of fake / real)
//   combine_kernel<D>:   S_pq -> Z -> loss, G -> H_pq -> C_pi    (L2 only)
//   sub_inv_kernel:      C_pi -> gradient sub-image pairs        (HBM write of grad)
// Only the REAL part of the inverse sub-image transforms reaches the gradient, so combine_kernel Hermitian-
// symmetrises each H_pq, Hs(k) = (H(k) + conj H(-k)) / 2, and packs two of them into one complex plane
// C_pi = Hs_{p,2i} + i Hs_{p,2i+1}: the inverse launch runs D^2/2 complex transforms whose real / imaginary
// outputs are the gradients of two horizontally adjacent pixels.
#pragma once
#include "line_tile.cuh"

namespace tfcfft {

struct SubCfg {
    static constexpr int LD = LineCfg::LD;
    static constexpr int NT_FWD = 128;  // two 64-thread groups, one sub-image of the pair each
    static constexpr int NT_INV = 64;
    static constexpr int LOAD_NI = 4;  // rows in flight per thread in the forward load (8 measured the same)
#ifndef TFC_STORE_NB
#define TFC_STORE_NB 8
#endif
    static constexpr int STORE_NB = TFC_STORE_NB;  // cluster inverse store: items whose tile reads are batched in front of their stores
    static constexpr size_t SMEM_FWD = (size_t)2 * 64 * LD * sizeof(float2);
    static constexpr size_t SMEM_INV = (size_t)64 * LD * sizeof(float2);
};

struct SubUnit {
    int tile_local, plane, p, i;  // tile within the chunk, pair plane, sub-image row phase, column-pair index
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
// workspace plane `plane` (float2[64][64]) of a chunk-local tile
TFC_HD float2* sub_plane(const Params& prm, int tile_local, int plane) {
    return reinterpret_cast<float2*>(prm.zws) + ((long long)tile_local * (prm.sub_d * prm.sub_d) + plane) * 4096;
}

// Workspace planes are written by one CTA and read by another within ONE launch of the pipelined kernel (and
// rewritten in between): reads bypass L1 (ld.global.cg), which may hold a stale copy of a line this SM read earlier.
TFC_HD float2 ws_load(const float2* p) {
#ifdef __CUDA_ARCH__
    return __ldcg(p);
#else
    return *p;
#endif
}

// ---- forward launch: sub-image pair (columns q = 2i, 2i+1 of row phase p) -> two complex work tiles ------
// The three loaders return whether every luma value of `fake` this THREAD folded equals the `real` one (the kernels
// reduce that over the tile: identical tiles get loss 0 and an exactly zero gradient, like the reference's L1Loss).
template <typename T, bool LUMA3, class Ctx>
TFC_HD bool sub_fwd_load(const Ctx& ctx, const Params& prm, const TileCoord& tc, const SubUnit& su, float2* s) {
    constexpr int LD = SubCfg::LD, NC = LUMA3 ? 3 : 1, NI = SubCfg::LOAD_NI;
    const int D = prm.sub_d, P = 64 * D;
    const T* fp = tile_ptr<T>(prm.fake, prm.fs, tc, P);
    const T* rp = real_tile_ptr<T>(prm, tc, P);
    const int fsh = (int)prm.fs[2], fsc = (int)prm.fs[1], rsh = (int)prm.rs[2], rsc = (int)prm.rs[1];
    const bool quant = (prm.flags & TFCFFT_QUANTIZE_U8) != 0;
    bool same = true;
    // one sub-image column b per lane (consecutive lanes -> consecutive 8-byte pixel pairs of the source row),
    // NI rows per thread in flight
#pragma unroll 1
    for (int it0 = ctx.tid; it0 < 4096; it0 += NI * ctx.nthreads) {
        float raw[NI][2][NC][2];  // [item][fake|real][channel][A|B]
#pragma unroll
        for (int u = 0; u < NI; ++u) {
            const int it = it0 + u * ctx.nthreads, b = it & 63, a = it >> 6;
            const int x = D * b + 2 * su.i, y = D * a + su.p;
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                IO<T>::load2(fp + y * fsh + c * fsc + x, raw[u][0][c]);
                IO<T>::load2(rp + y * rsh + c * rsc + x, raw[u][1][c]);
            }
        }
#pragma unroll
        for (int u = 0; u < NI; ++u) {
            const int it = it0 + u * ctx.nthreads, b = it & 63, a = it >> 6;
            float z[2][2];  // [fake|real][A|B]
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int l = 0; l < 2; ++l) {
                    if (!quant) {
                        float f = prm.lw[0] * raw[u][h][0][l];
                        if constexpr (LUMA3) f = fmaf(prm.lw[2], raw[u][h][2][l], fmaf(prm.lw[1], raw[u][h][1][l], f));
                        z[h][l] = f;
                    } else if constexpr (LUMA3) {
                        z[h][l] = (float)((19595 * IO<T>::quant(raw[u][h][0][l]) + 38470 * IO<T>::quant(raw[u][h][1][l]) +
                                           7471 * IO<T>::quant(raw[u][h][2][l]) + 0x8000) >> 16);
                    } else {
                        z[h][l] = (float)IO<T>::quant(raw[u][h][0][l]);
                    }
                }
            s[a * LD + b] = make_float2(z[0][0], z[1][0]);
            s[64 * LD + a * LD + b] = make_float2(z[0][1], z[1][1]);
            same = same && z[0][0] == z[1][0] && z[0][1] == z[1][1];
        }
    }
    return same;
}
// A/B builds (-DTFCFFT_D2_PAIR8): D = 2 on the generic 8-byte forms
TFC_HD bool sub_d2_quads(const Params& prm) {
#ifdef TFCFFT_D2_PAIR8
    (void)prm;
    return false;
#else
    return prm.sub_d == 2;
#endif
}
// D = 2 variant (one unit = one row phase: both column phases of every second row): 16-byte loads -- two pixel pairs
// (A_b, B_b, A_b+1, B_b+1) per lane, half the load instructions of the 8-byte form (ncu: 17 % of the launch was load issue)
template <typename T, bool LUMA3, class Ctx>
TFC_HD bool sub_fwd_load_d2(const Ctx& ctx, const Params& prm, const TileCoord& tc, const SubUnit& su, float2* s) {
    constexpr int LD = SubCfg::LD, NC = LUMA3 ? 3 : 1, NI = SubCfg::LOAD_NI, P = 128;
    const T* fp = tile_ptr<T>(prm.fake, prm.fs, tc, P);
    const T* rp = real_tile_ptr<T>(prm, tc, P);
    const int fsh = (int)prm.fs[2], fsc = (int)prm.fs[1], rsh = (int)prm.rs[2], rsc = (int)prm.rs[1];
    const bool quant = (prm.flags & TFCFFT_QUANTIZE_U8) != 0;
    bool same = true;
#pragma unroll 1
    for (int it0 = ctx.tid; it0 < 2048; it0 += NI * ctx.nthreads) {
        float raw[NI][2][NC][4];  // [item][fake|real][channel][A_b, B_b, A_b+1, B_b+1]
#pragma unroll
        for (int u = 0; u < NI; ++u) {
            const int it = it0 + u * ctx.nthreads, bp = it & 31, a = it >> 5;
            const int x = 4 * bp, y = 2 * a + su.p;
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                IO<T>::load4(fp + y * fsh + c * fsc + x, raw[u][0][c]);
                IO<T>::load4(rp + y * rsh + c * rsc + x, raw[u][1][c]);
            }
        }
#pragma unroll
        for (int u = 0; u < NI; ++u) {
            const int it = it0 + u * ctx.nthreads, bp = it & 31, a = it >> 5;
            float z[2][4];  // [fake|real][A_b, B_b, A_b+1, B_b+1]
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int l = 0; l < 4; ++l) {
                    if (!quant) {
                        float f = prm.lw[0] * raw[u][h][0][l];
                        if constexpr (LUMA3) f = fmaf(prm.lw[2], raw[u][h][2][l], fmaf(prm.lw[1], raw[u][h][1][l], f));
                        z[h][l] = f;
                    } else if constexpr (LUMA3) {
                        z[h][l] = (float)((19595 * IO<T>::quant(raw[u][h][0][l]) + 38470 * IO<T>::quant(raw[u][h][1][l]) +
                                           7471 * IO<T>::quant(raw[u][h][2][l]) + 0x8000) >> 16);
                    } else {
                        z[h][l] = (float)IO<T>::quant(raw[u][h][0][l]);
                    }
                }
            float2* d = s + a * LD + 2 * bp;
            d[0] = make_float2(z[0][0], z[1][0]);
            d[1] = make_float2(z[0][2], z[1][2]);
            d[64 * LD] = make_float2(z[0][1], z[1][1]);
            d[64 * LD + 1] = make_float2(z[0][3], z[1][3]);
            same = same && z[0][0] == z[1][0] && z[0][1] == z[1][1] && z[0][2] == z[1][2] && z[0][3] == z[1][3];
        }
    }
    return same;
}
// D = 4 variant for a 2-CTA cluster (one cluster = one row phase p of a tile = the two column pairs i = 0, 1): each
// CTA reads HALF of the rows with full 16-byte loads (all four column phases q: every 32-byte sector is fetched
// once instead of twice) and distributes the luma values: q = 0, 1 -> the work tiles of the i = 0 CTA (dst01),
// q = 2, 3 -> those of the i = 1 CTA (dst23); one of the two destinations is the peer CTA's shared memory.
template <typename T, bool LUMA3, class Ctx>
TFC_HD bool sub_fwd_load_quad(const Ctx& ctx, const Params& prm, const TileCoord& tc, int p, int half, float2* dst01, float2* dst23) {
    constexpr int LD = SubCfg::LD, NC = LUMA3 ? 3 : 1, NI = SubCfg::LOAD_NI, P = 256;
    const T* fp = tile_ptr<T>(prm.fake, prm.fs, tc, P);
    const T* rp = real_tile_ptr<T>(prm, tc, P);
    const int fsh = (int)prm.fs[2], fsc = (int)prm.fs[1], rsh = (int)prm.rs[2], rsc = (int)prm.rs[1];
    const bool quant = (prm.flags & TFCFFT_QUANTIZE_U8) != 0;
    bool same = true;
#pragma unroll 1
    for (int it0 = ctx.tid; it0 < 2048; it0 += NI * ctx.nthreads) {
        float raw[NI][2][NC][4];  // [item][fake|real][channel][q]
#pragma unroll
        for (int u = 0; u < NI; ++u) {
            const int it = it0 + u * ctx.nthreads, b = it & 63, a = 32 * half + (it >> 6);
            const int x = 4 * b, y = 4 * a + p;
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                IO<T>::load4(fp + y * fsh + c * fsc + x, raw[u][0][c]);
                IO<T>::load4(rp + y * rsh + c * rsc + x, raw[u][1][c]);
            }
        }
#pragma unroll
        for (int u = 0; u < NI; ++u) {
            const int it = it0 + u * ctx.nthreads, b = it & 63, a = 32 * half + (it >> 6);
            float z[2][4];  // [fake|real][q]
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (!quant) {
                        float f = prm.lw[0] * raw[u][h][0][q];
                        if constexpr (LUMA3) f = fmaf(prm.lw[2], raw[u][h][2][q], fmaf(prm.lw[1], raw[u][h][1][q], f));
                        z[h][q] = f;
                    } else if constexpr (LUMA3) {
                        z[h][q] = (float)((19595 * IO<T>::quant(raw[u][h][0][q]) + 38470 * IO<T>::quant(raw[u][h][1][q]) +
                                           7471 * IO<T>::quant(raw[u][h][2][q]) + 0x8000) >> 16);
                    } else {
                        z[h][q] = (float)IO<T>::quant(raw[u][h][0][q]);
                    }
                }
            dst01[a * LD + b] = make_float2(z[0][0], z[1][0]);
            dst01[64 * LD + a * LD + b] = make_float2(z[0][1], z[1][1]);
            dst23[a * LD + b] = make_float2(z[0][2], z[1][2]);
            dst23[64 * LD + a * LD + b] = make_float2(z[0][3], z[1][3]);
            same = same && z[0][0] == z[1][0] && z[0][1] == z[1][1] && z[0][2] == z[1][2] && z[0][3] == z[1][3];
        }
    }
    return same;
}

// D = 8 variant for a 4-CTA cluster (one cluster = one row phase p of a 512 x 512 tile = the four column pairs
// i = 0..3): each CTA reads a QUARTER of the rows with two 16-byte loads per pixel octet (all eight column phases q:
// every 32-byte sector is fetched once instead of four times) and distributes the luma values: q = 2i, 2i+1 -> the
// work tiles of CTA i (dst[i]); three of the four destinations are peer CTAs' shared memory.
template <typename T, bool LUMA3, class Ctx>
TFC_HD bool sub_fwd_load_oct(const Ctx& ctx, const Params& prm, const TileCoord& tc, int p, int quarter, float2* const (&dst)[4]) {
    constexpr int LD = SubCfg::LD, NC = LUMA3 ? 3 : 1, NI = 2, P = 512;
    const T* fp = tile_ptr<T>(prm.fake, prm.fs, tc, P);
    const T* rp = real_tile_ptr<T>(prm, tc, P);
    const int fsh = (int)prm.fs[2], fsc = (int)prm.fs[1], rsh = (int)prm.rs[2], rsc = (int)prm.rs[1];
    const bool quant = (prm.flags & TFCFFT_QUANTIZE_U8) != 0;
    bool same = true;
#pragma unroll 1
    for (int it0 = ctx.tid; it0 < 1024; it0 += NI * ctx.nthreads) {
        float raw[NI][2][NC][8];  // [item][fake|real][channel][q]
#pragma unroll
        for (int u = 0; u < NI; ++u) {
            const int it = it0 + u * ctx.nthreads, b = it & 63, a = 16 * quarter + (it >> 6);
            const int x = 8 * b, y = 8 * a + p;
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                IO<T>::load4(fp + y * fsh + c * fsc + x, raw[u][0][c]);
                IO<T>::load4(fp + y * fsh + c * fsc + x + 4, raw[u][0][c] + 4);
                IO<T>::load4(rp + y * rsh + c * rsc + x, raw[u][1][c]);
                IO<T>::load4(rp + y * rsh + c * rsc + x + 4, raw[u][1][c] + 4);
            }
        }
#pragma unroll
        for (int u = 0; u < NI; ++u) {
            const int it = it0 + u * ctx.nthreads, b = it & 63, a = 16 * quarter + (it >> 6);
            float z[2][8];  // [fake|real][q]
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    if (!quant) {
                        float f = prm.lw[0] * raw[u][h][0][q];
                        if constexpr (LUMA3) f = fmaf(prm.lw[2], raw[u][h][2][q], fmaf(prm.lw[1], raw[u][h][1][q], f));
                        z[h][q] = f;
                    } else if constexpr (LUMA3) {
                        z[h][q] = (float)((19595 * IO<T>::quant(raw[u][h][0][q]) + 38470 * IO<T>::quant(raw[u][h][1][q]) +
                                           7471 * IO<T>::quant(raw[u][h][2][q]) + 0x8000) >> 16);
                    } else {
                        z[h][q] = (float)IO<T>::quant(raw[u][h][0][q]);
                    }
                }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                dst[i][a * LD + b] = make_float2(z[0][2 * i], z[1][2 * i]);
                dst[i][64 * LD + a * LD + b] = make_float2(z[0][2 * i + 1], z[1][2 * i + 1]);
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) same = same && z[0][q] == z[1][q];
        }
    }
    return same;
}

// Forward passes of both work tiles on ONE copy of the 64-point core (callers keep the pass loop rolled, like
// line_fft_pass): pass 0 = rows (one thread per row, in place), pass 1 = columns, written straight to the workspace
// planes (coalesced across the column index).
template <class Ctx>
TFC_HD void sub_fwd_pass(const Ctx& ctx, const Params& prm, const SubUnit& su, float2* s, int pass) {
    constexpr int LD = SubCfg::LD;
    for (int l = ctx.tid; l < 128; l += ctx.nthreads) {
        const int t = l >> 6, x = l & 63;
        float2* row = s + l * LD;  // tile (l >> 6), row (l & 63): the two tiles are contiguous
        const float2* col = s + t * 64 * LD + x;
        float2 v[64];
        if (pass == 0) {
#pragma unroll
            for (int i = 0; i < 64; ++i) v[i] = row[i];
        } else {
#pragma unroll
            for (int y = 0; y < 64; ++y) v[y] = col[y * LD];
        }
        fft64<false>(v);
        if (pass == 0) {
#pragma unroll
            for (int sl = 0; sl < 64; ++sl) row[fft64_freq(sl)] = v[sl];
        } else {
            float2* plane = sub_plane(prm, su.tile_local, su.p * prm.sub_d + 2 * su.i + t) + x;
#pragma unroll
            for (int sl = 0; sl < 64; ++sl) plane[fft64_freq(sl) * 64] = v[sl];
        }
    }
}
template <class Ctx>
TFC_HD void sub_fwd_rows(const Ctx& ctx, float2* s) {
    sub_fwd_pass(ctx, Params{}, SubUnit{}, s, 0);
}
template <class Ctx>
TFC_HD void sub_fwd_cols_store(const Ctx& ctx, const Params& prm, const SubUnit& su, float2* s) {
    sub_fwd_pass(ctx, prm, su, s, 1);
}

// ---- inverse launch: packed plane C_pi -> gradients of the sub-image pair ----------------------------
// Inverse passes on ONE copy of the forward core: IDFT(x)[n] = DFT(x)[(64 - n) mod 64], so the results are stored at
// the mirrored index.  pass 0 = columns straight from the workspace plane (64 independent 8-byte loads in flight per
// thread), pass 1 = rows in shared memory (real / imaginary parts = gradients of two horizontally adjacent pixels).
template <class Ctx>
TFC_HD void sub_inv_pass(const Ctx& ctx, const Params& prm, const SubUnit& su, float2* s, int pass) {
    constexpr int LD = SubCfg::LD;
    for (int l = ctx.tid; l < 64; l += ctx.nthreads) {
        float2* row = s + l * LD;
        float2 v[64];
        if (pass == 0) {
            const float2* plane = sub_plane(prm, su.tile_local, su.plane) + l;
#pragma unroll
            for (int y = 0; y < 64; ++y) v[y] = ws_load(plane + y * 64);
        } else {
#pragma unroll
            for (int k = 0; k < 64; ++k) v[k] = row[k];
        }
        fft64<false>(v);
        if (pass == 0) {
#pragma unroll
            for (int sl = 0; sl < 64; ++sl) s[((64 - fft64_freq(sl)) & 63) * LD + l] = v[sl];
        } else {
#pragma unroll
            for (int sl = 0; sl < 64; ++sl) row[(64 - fft64_freq(sl)) & 63] = v[sl];  // (grad of pixel 2i, grad of pixel 2i+1)
        }
    }
}
template <class Ctx>
TFC_HD void sub_inv_cols(const Ctx& ctx, const Params& prm, const SubUnit& su, float2* s) {
    sub_inv_pass(ctx, prm, su, s, 0);
}
template <class Ctx>
TFC_HD void sub_inv_rows(const Ctx& ctx, float2* s) {
    sub_inv_pass(ctx, Params{}, SubUnit{}, s, 1);
}
template <typename T, bool LUMA3, class Ctx>
TFC_HD void sub_inv_store(const Ctx& ctx, const Params& prm, const TileCoord& tc, const SubUnit& su, const float2* s) {
    constexpr int LD = SubCfg::LD, NC = LUMA3 ? 3 : 1;
    const int D = prm.sub_d, P = 64 * D;
    T* gp = const_cast<T*>(tile_ptr<T>(prm.grad, prm.gs, tc, P));
    const int sh = (int)prm.gs[2], sc = (int)prm.gs[1];
    const GradOut go = grad_out(prm);
#pragma unroll 4
    for (int it = ctx.tid; it < 4096; it += ctx.nthreads) {
        const int b = it & 63, a = it >> 6;  // consecutive lanes: consecutive 8-byte pairs of the gradient row
        const int x = D * b + 2 * su.i, y = D * a + su.p;
        const float2 g = s[a * LD + b];
#pragma unroll
        for (int c = 0; c < NC; ++c) grad_store2<T>(go, gp + y * sh + c * sc + x, go.w[c] * g.x, go.w[c] * g.y);
    }
}

// D = 2 variant: 16-byte stores -- the gradients of two adjacent pixel pairs per lane (half the store instructions and
// address arithmetic of the 8-byte form; ncu: the store loop was 53 % of the launch, issue-bound)
template <typename T, bool LUMA3, class Ctx>
TFC_HD void sub_inv_store_d2(const Ctx& ctx, const Params& prm, const TileCoord& tc, const SubUnit& su, const float2* s) {
    constexpr int LD = SubCfg::LD, NC = LUMA3 ? 3 : 1, P = 128;
    T* gp = const_cast<T*>(tile_ptr<T>(prm.grad, prm.gs, tc, P));
    const int sh = (int)prm.gs[2], sc = (int)prm.gs[1];
    const GradOut go = grad_out(prm);
#pragma unroll 4
    for (int it = ctx.tid; it < 2048; it += ctx.nthreads) {
        const int bp = it & 31, a = it >> 5;  // consecutive lanes: consecutive 16-byte quads of the gradient row
        const int x = 4 * bp, y = 2 * a + su.p;
        const float2 g0 = s[a * LD + 2 * bp], g1 = s[a * LD + 2 * bp + 1];
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            float v[4] = {go.w[c] * g0.x, go.w[c] * g0.y, go.w[c] * g1.x, go.w[c] * g1.y};
            grad_store4<T>(go, gp + y * sh + c * sc + x, v);
        }
    }
}

// D = 4 variant for a 2-CTA cluster (mirror of sub_fwd_load_quad): the CTA writes HALF of the gradient rows of row
// phase p with full 16-byte stores, taking pixels q = 0, 1 from the i = 0 CTA's tile (s01) and q = 2, 3 from the
// i = 1 CTA's (s23); one of the two is the peer CTA's shared memory.
template <typename T, bool LUMA3, class Ctx>
TFC_HD void sub_inv_store_quad(const Ctx& ctx, const Params& prm, const TileCoord& tc, int p, int half, const float2* own,
                               const float2* peer) {
    constexpr int LD = SubCfg::LD, NC = LUMA3 ? 3 : 1, P = 256;
    T* gp = const_cast<T*>(tile_ptr<T>(prm.grad, prm.gs, tc, P));
    const int sh = (int)prm.gs[2], sc = (int)prm.gs[1];
    const GradOut go = grad_out(prm);
    // `peer` is the other CTA's tile, read through a generic pointer; the compiler cannot move such loads across the
    // global stores, so a plainly unrolled loop waits for one distributed-shared-memory round trip PER ITEM (ncu: 56 % of
    // the launch's stall samples on the first use of these loads).  All peer values of a batch of NB items are
    // requested in front of the batch's stores; the CTA's own tile is read as plain shared memory next to them.
    constexpr int NB = SubCfg::STORE_NB;
#pragma unroll 1
    for (int it0 = ctx.tid; it0 < 2048; it0 += NB * ctx.nthreads) {
        float2 rem[NB];
#pragma unroll
        for (int u = 0; u < NB; ++u) rem[u] = peer[(32 * half + ((it0 + u * ctx.nthreads) >> 6)) * LD + ((it0 + u * ctx.nthreads) & 63)];
#pragma unroll
        for (int u = 0; u < NB; ++u) {
            const int it = it0 + u * ctx.nthreads, bb = it & 63, a = 32 * half + (it >> 6);
            const float2 loc = own[a * LD + bb];
            const float2 lo = half == 0 ? loc : rem[u], hi = half == 0 ? rem[u] : loc;
            const int x = 4 * bb, y = 4 * a + p;
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                float v[4] = {go.w[c] * lo.x, go.w[c] * lo.y, go.w[c] * hi.x, go.w[c] * hi.y};
                grad_store4<T>(go, gp + y * sh + c * sc + x, v);
            }
        }
    }
}

// D = 4 without clusters: one CTA holds BOTH packed planes of a row phase (s01: pixels q = 0, 1; s23: q = 2, 3; thanks
// to the Hermitian packing that is 66 KB, the size of one forward work-tile pair), so all four column phases are
// local: full 16-byte stores straight from plain shared memory, no distributed-shared-memory traffic at all.
template <typename T, bool LUMA3, class Ctx>
TFC_HD void sub_inv_store_rows4(const Ctx& ctx, const Params& prm, const TileCoord& tc, int p, const float2* s01, const float2* s23) {
    constexpr int LD = SubCfg::LD, NC = LUMA3 ? 3 : 1, P = 256;
    T* gp = const_cast<T*>(tile_ptr<T>(prm.grad, prm.gs, tc, P));
    const int sh = (int)prm.gs[2], sc = (int)prm.gs[1];
    const GradOut go = grad_out(prm);
#pragma unroll 4
    for (int it = ctx.tid; it < 4096; it += ctx.nthreads) {
        const int b = it & 63, a = it >> 6;
        const float2 lo = s01[a * LD + b], hi = s23[a * LD + b];
        const int x = 4 * b, y = 4 * a + p;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            float v[4] = {go.w[c] * lo.x, go.w[c] * lo.y, go.w[c] * hi.x, go.w[c] * hi.y};
            grad_store4<T>(go, gp + y * sh + c * sc + x, v);
        }
    }
}

// D = 8 variant for a 4-CTA cluster (mirror of sub_fwd_load_oct): the CTA writes a QUARTER of the gradient rows of
// row phase p with two 16-byte stores per pixel octet, taking pixels q = 2i, 2i+1 from CTA i's tile (src[i]).
template <typename T, bool LUMA3, class Ctx>
TFC_HD void sub_inv_store_oct(const Ctx& ctx, const Params& prm, const TileCoord& tc, int p, int quarter, const float2* const (&src)[4]) {
    constexpr int LD = SubCfg::LD, NC = LUMA3 ? 3 : 1, P = 512;
    T* gp = const_cast<T*>(tile_ptr<T>(prm.grad, prm.gs, tc, P));
    const int sh = (int)prm.gs[2], sc = (int)prm.gs[1];
    const GradOut go = grad_out(prm);
    // loads of NB items in one batch in front of their stores (see sub_inv_store_quad)
    constexpr int NB = SubCfg::STORE_NB >= 8 ? 4 : 1;
#pragma unroll 1
    for (int it0 = ctx.tid; it0 < 1024; it0 += NB * ctx.nthreads) {
        float2 g[NB][4];
#pragma unroll
        for (int u = 0; u < NB; ++u) {
            const int it = it0 + u * ctx.nthreads, b = it & 63, a = 16 * quarter + (it >> 6);
#pragma unroll
            for (int i = 0; i < 4; ++i) g[u][i] = src[i][a * LD + b];
        }
#pragma unroll
        for (int u = 0; u < NB; ++u) {
            const int it = it0 + u * ctx.nthreads, b = it & 63, a = 16 * quarter + (it >> 6);
            const int x = 8 * b, y = 8 * a + p;
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                float v0[4] = {go.w[c] * g[u][0].x, go.w[c] * g[u][0].y, go.w[c] * g[u][1].x, go.w[c] * g[u][1].y};
                float v1[4] = {go.w[c] * g[u][2].x, go.w[c] * g[u][2].y, go.w[c] * g[u][3].x, go.w[c] * g[u][3].y};
                grad_store4<T>(go, gp + y * sh + c * sc + x, v0);
                grad_store4<T>(go, gp + y * sh + c * sc + x + 4, v1);
            }
        }
    }
}

template <typename T, bool LUMA3, class Ctx>
TFC_HD bool sub_fwd_process(const Ctx& ctx, const Params& prm, int u, float2* s) {
    const SubUnit su = sub_unit(u, prm.sub_d);
    ctx.mark(0);
    const TileCoord tc = decode_tile(prm, prm.tile_base + su.tile_local);
    const bool same = sub_d2_quads(prm) ? sub_fwd_load_d2<T, LUMA3>(ctx, prm, tc, su, s) : sub_fwd_load<T, LUMA3>(ctx, prm, tc, su, s);
    ctx.sync();
    ctx.mark(1);
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {  // rolled: one copy of the 64-point core
        sub_fwd_pass(ctx, prm, su, s, pass);
        ctx.sync();
        ctx.mark(2 + pass);
    }
    return same;
}
template <typename T, bool LUMA3, class Ctx>
TFC_HD void sub_inv_process(const Ctx& ctx, const Params& prm, int u, float2* s) {
    const SubUnit su = sub_unit(u, prm.sub_d);
    ctx.mark(0);
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {  // rolled: one copy of the 64-point core
        sub_inv_pass(ctx, prm, su, s, pass);
        ctx.sync();
        ctx.mark(1 + pass);
    }
    const TileCoord tc = decode_tile(prm, prm.tile_base + su.tile_local);
    if (sub_d2_quads(prm)) sub_inv_store_d2<T, LUMA3>(ctx, prm, tc, su, s);
    else sub_inv_store<T, LUMA3>(ctx, prm, tc, su, s);
    ctx.sync();
    ctx.mark(3);
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

// position pairs {(ky',kx'), -(ky',kx')} of the 64 x 64 sub-grid: 64 x 32 slots walk the half-plane columns
// kx' = 1..31 in memory order (slot j = 0 is idle: the kx' = 0 column is handled, with kx' = 32, by the 66
// items that pair rows instead)
// Out-of-line copy of the packed bin evaluation: combine_item calls it from 8 (D = 4) fully unrolled entry
// pairs; inlining it there made the kernel instruction-fetch bound (ncu: no_instruction 3.9 per issue).
#ifdef __CUDA_ARCH__
__device__ __noinline__
#else
inline
#endif
c2 bin_eval_pair_call(const Params& prm, bool mse, bool phase, c2 zk, c2 zm, float2& accA, float2& accP) {
    return bin_eval_pair(prm, mse, phase, zk, zm, accA, accP);  // rare path (self-conjugate columns)
}

// two independent packed evaluations per call; everything by value (register ABI: no local-memory traffic)
struct QuadEval {
    c2 g0, g1;
    float2 a, p;  // loss-term increments
};
#ifdef __CUDA_ARCH__
__device__ __noinline__
#else
inline
#endif
QuadEval bin_eval_quad_call(const Params& prm, bool mse, bool phase, c2 zk0, c2 zm0, c2 zk1, c2 zm1) {
    QuadEval r;
    float2 a0 = make_float2(0.f, 0.f), p0 = a0, a1 = a0, p1 = a0;  // separate accumulators: independent chains
    r.g0 = bin_eval_pair(prm, mse, phase, zk0, zm0, a0, p0);
    r.g1 = bin_eval_pair(prm, mse, phase, zk1, zm1, a1, p1);
    r.a = p_add(a0, a1);
    r.p = p_add(p0, p1);
    return r;
}

// four independent packed evaluations per call (A/B: -DTFC_EVAL_OCT)
struct OctEval {
    c2 g[4];
    float2 a, p;
};
#ifdef __CUDA_ARCH__
__device__ __noinline__
#else
inline
#endif
OctEval bin_eval_oct_call(const Params& prm, bool mse, bool phase, c2 zk0, c2 zm0, c2 zk1, c2 zm1, c2 zk2, c2 zm2, c2 zk3, c2 zm3) {
    OctEval r;
    const float2 z = make_float2(0.f, 0.f);
    float2 a0 = z, p0 = z, a1 = z, p1 = z, a2 = z, p2 = z, a3 = z, p3 = z;
    r.g[0] = bin_eval_pair(prm, mse, phase, zk0, zm0, a0, p0);
    r.g[1] = bin_eval_pair(prm, mse, phase, zk1, zm1, a1, p1);
    r.g[2] = bin_eval_pair(prm, mse, phase, zk2, zm2, a2, p2);
    r.g[3] = bin_eval_pair(prm, mse, phase, zk3, zm3, a3, p3);
    r.a = p_add(p_add(a0, a1), p_add(a2, a3));
    r.p = p_add(p_add(p0, p1), p_add(p2, p3));
    return r;
}

#ifdef __CUDA_ARCH__
__device__ __noinline__
#else
inline
#endif
float2 bin_eval_call(const Params& prm, float2 zk, float2 zm, float mult, float& accA, float& accP) {
    return bin_eval(prm, zk, zm, mult, accA, accP);
}

// One position pair of one tile: load the D^2 sub-spectra at both positions, combine, evaluate every
// full-size half-plane bin they contain, un-combine the spectral gradient, store back.  `same`: the tile's fake and
// real pixels are identical -- every loss term and its gradient vanish exactly (sign(0) = 0 in the reference's
// L1Loss), so the packed planes of the inverse launch are zero-filled instead.
// `eqf`: the tile's flag bytes (nullptr: not tracked), D*D/2 of them, naturally aligned
template <int D>
TFC_HD void combine_item(const Params& prm, float2* ws_tile, int item, float& accA, float& accP, const unsigned char* eqf = nullptr) {
    constexpr int P = 64 * D, HD = D / 2;
    int kyA, kxA;
    if (item < 64 * 32) {
        kxA = item & 31;  // 1 .. 31
        if (kxA == 0) return;
        kyA = item >> 5;
    } else {
        const int sp = item - 64 * 32;
        kxA = (sp / 33) * 32;
        kyA = sp % 33;
    }
    const int kyB = (64 - kyA) & 63, kxB = (64 - kxA) & 63;
    const bool self = (kyA == kyB) && (kxA == kxB);
    const int offA = kyA * 64 + kxA, offB = kyB * 64 + kxB;
    // the flag word is loaded first and tested AFTER the sub-spectra loads are in flight: its latency hides under them
    static_assert(D == 2 || D == 4, "D = 8 runs combine8_rows");
    unsigned long long flags = 0ull;
    if (eqf != nullptr) {
#ifdef __CUDA_ARCH__
        if constexpr (D == 2) flags = __ldcg(reinterpret_cast<const unsigned short*>(eqf));
        else flags = __ldcg(reinterpret_cast<const unsigned long long*>(eqf));
#else
        for (int i = 0; i < D * D / 2; ++i) flags |= (unsigned long long)eqf[i] << (8 * i);
#endif
    }
    float2 za[D][D], zb[D][D];
#pragma unroll
    for (int p = 0; p < D; ++p)
#pragma unroll
        for (int q = 0; q < D; ++q) {
            za[p][q] = ws_load(ws_tile + (p * D + q) * 4096 + offA);
            zb[p][q] = ws_load(ws_tile + (p * D + q) * 4096 + offB);
        }
    const bool same = flags == (D == 2 ? 0x0101ull : 0x0101010101010101ull);
    if (same) {
        if (prm.grad != nullptr) {
            const float2 z0 = make_float2(0.f, 0.f);
#pragma unroll
            for (int pl = 0; pl < D * HD; ++pl) {
                ws_tile[pl * 4096 + offA] = z0;
                if (!self) ws_tile[pl * 4096 + offB] = z0;
            }
        }
        return;
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
        // default modes: two entries (be, be+1) per packed evaluation, two packed evaluations per out-of-line call
        // (independent dependency chains: the call is latency-bound on MUFU / polynomial chains otherwise)
        const bool mse = (prm.flags & TFCFFT_DIST_MSE) != 0, phase = !(prm.flags & TFCFFT_NO_PHASE);
        float2 pA = make_float2(0.f, 0.f), pP = make_float2(0.f, 0.f);
        const float2 z0 = make_float2(0.f, 0.f);
        constexpr int NE = D * D / 2;  // packed evaluations: entry pairs (al, 2 bp), (al, 2 bp + 1)
#ifdef TFC_EVAL_OCT
        constexpr int EV = NE >= 4 ? 4 : 2;  // packed evaluations per out-of-line call
#else
        constexpr int EV = 2;
#endif
#pragma unroll
        for (int e0 = 0; e0 < NE; e0 += EV) {
            bool isM[EV][2], both[EV][2], live[EV][2];
            float2 k_[EV][2], m_[EV][2];
            c2 zk[EV], zm[EV], g[EV];
#pragma unroll
            for (int h = 0; h < EV; ++h) {
                const int al = (e0 + h) / HD, bp = (e0 + h) % HD;
                const int alB = kyA ? D - 1 - al : (D - al) % D;
#pragma unroll
                for (int l = 0; l < 2; ++l) {
                    const int be = 2 * bp + l;
                    const int beB = kxA ? D - 1 - be : (D - be) % D;
                    const int kxf = kxA + 64 * be;
                    live[h][l] = !(self && (al * D + be) > (alB * D + beB));
                    const bool special = (kxf == 0 || kxf == P / 2);
                    both[h][l] = live[h][l] && special && !(self && al == alB && be == beB);
                    isM[h][l] = !special && kxf > P / 2;
                    k_[h][l] = live[h][l] ? (isM[h][l] ? zp[al][be] : za[al][be]) : z0;
                    m_[h][l] = live[h][l] ? (isM[h][l] ? za[al][be] : zp[al][be]) : z0;
                }
                zk[h] = make_c2(make_float2(k_[h][0].x, k_[h][1].x), make_float2(k_[h][0].y, k_[h][1].y));
                zm[h] = make_c2(make_float2(m_[h][0].x, m_[h][1].x), make_float2(m_[h][0].y, m_[h][1].y));
            }
            if constexpr (EV == 4) {
                const OctEval q = bin_eval_oct_call(prm, mse, phase, zk[0], zm[0], zk[1], zm[1], zk[2], zm[2], zk[3], zm[3]);
#pragma unroll
                for (int h = 0; h < 4; ++h) g[h] = q.g[h];
                pA = p_add(pA, q.a);
                pP = p_add(pP, q.p);
            } else {
                const QuadEval q = bin_eval_quad_call(prm, mse, phase, zk[0], zm[0], zk[1], zm[1]);
                g[0] = q.g0;
                g[1] = q.g1;
                pA = p_add(pA, q.a);
                pP = p_add(pP, q.p);
            }
#pragma unroll
            for (int h = 0; h < EV; ++h) {
                const int al = (e0 + h) / HD, bp = (e0 + h) % HD;
                c2 g2 = make_c2(z0, z0);
                if (both[h][0] || both[h][1]) {  // self-conjugate columns only: evaluate the mirrored bin as well
                    const c2 zk2 = make_c2(make_float2(both[h][0] ? m_[h][0].x : 0.f, both[h][1] ? m_[h][1].x : 0.f),
                                           make_float2(both[h][0] ? m_[h][0].y : 0.f, both[h][1] ? m_[h][1].y : 0.f));
                    const c2 zm2 = make_c2(make_float2(both[h][0] ? k_[h][0].x : 0.f, both[h][1] ? k_[h][1].x : 0.f),
                                           make_float2(both[h][0] ? k_[h][0].y : 0.f, both[h][1] ? k_[h][1].y : 0.f));
                    g2 = bin_eval_pair_call(prm, mse, phase, zk2, zm2, pA, pP);
                }
#pragma unroll
                for (int l = 0; l < 2; ++l) {
                    const int be = 2 * bp + l;
                    const float2 gl = l ? make_float2(g[h].re.y, g[h].im.y) : make_float2(g[h].re.x, g[h].im.x);
                    const float2 g2l = l ? make_float2(g2.re.y, g2.im.y) : make_float2(g2.re.x, g2.im.x);
                    ga[al][be] = live[h][l] ? (isM[h][l] ? z0 : gl) : z0;
                    gp[al][be] = live[h][l] ? (isM[h][l] ? gl : (both[h][l] ? g2l : z0)) : z0;
                }
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
    // Hermitian-symmetrise each H_pq over the position pair and pack two of them per complex plane (in place:
    // this item owns positions A and B of every plane)
#pragma unroll
    for (int p = 0; p < D; ++p)
#pragma unroll
        for (int i = 0; i < HD; ++i) {
            float2 h0, h1;  // Hs_{p,2i}(A), Hs_{p,2i+1}(A); Hs(B) = conj Hs(A)
            if (self) {
                h0 = make_float2(ga[p][2 * i].x, 0.f);
                h1 = make_float2(ga[p][2 * i + 1].x, 0.f);
            } else {
                h0 = make_float2(0.5f * (ga[p][2 * i].x + gb[p][2 * i].x), 0.5f * (ga[p][2 * i].y - gb[p][2 * i].y));
                h1 = make_float2(0.5f * (ga[p][2 * i + 1].x + gb[p][2 * i + 1].x), 0.5f * (ga[p][2 * i + 1].y - gb[p][2 * i + 1].y));
            }
            float2* plane = ws_tile + (p * HD + i) * 4096;
            plane[offA] = make_float2(h0.x - h1.y, h0.y + h1.x);
            if (!self) plane[offB] = make_float2(h0.x + h1.y, h1.x - h0.y);
        }
}

TFC_HD bool sub_supported(const Params& prm) {
    return (prm.p == 128 || prm.p == 256 || prm.p == 512) && prm.spec_mode == 0 && !(prm.flags & (TFCFFT_FORCE_SPLIT | TFCFFT_FORCE_GENERIC));
}

}  // namespace tfcfft
