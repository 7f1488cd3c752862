// emu_fft.cu -- serial CPU execution of the FFT-loss kernels' arithmetic for ONE element type (-DTFC_DT=0..3); part of
// libtfcfft_emu.so (TEST INFRASTRUCTURE, see emu.cu).  Split per dtype so that the host compile parallelises.
#define TFCFFT_EMU_BUILD 1  // serial emulation: the one-thread combine item (9 partial sums per 256 x 256 tile)
#include <cstdlib>
#include <vector>

#include "launchers.h"
#include "pair_tile.cuh"
#include "sub_tile.cuh"
#include "combine8.cuh"
#include "line_tile.cuh"

using namespace tfcfft;

namespace {

template <int P, typename T, bool LUMA3>
void run_resident(Params prm) {
    SerialCtx ctx;
    std::vector<float2> s((size_t)P * (P + 1)), tw(P);
    fill_twiddles<P>(ctx, tw.data());
    for (int tile = 0; tile < prm.tiles_total; ++tile) {
        float a = 0.f, p = 0.f;
        tile_process<P, T, LUMA3>(ctx, prm, tile, s.data(), tw.data(), a, p);
        prm.partials[2 * tile] = a;
        prm.partials[2 * tile + 1] = p;
    }
}

template <typename T, bool LUMA3>
void run_line(Params prm) {
    SerialCtx ctx;
    std::vector<float2> s((size_t)64 * LineCfg::LD);
    for (int tile = 0; tile < prm.tiles_total; ++tile) {
        float a = 0.f, p = 0.f;
        line_process<T, LUMA3>(ctx, prm, tile, s.data(), a, p, -1, (prm.flags & TFCFFT_USE_HALFLINE) != 0);
        prm.partials[2 * tile] = a;
        prm.partials[2 * tile + 1] = p;
    }
}

template <int P, typename T, bool LUMA3>
void run_pair(Params prm) {
    if constexpr (P == 64) {
        SerialCtx ctx;
        std::vector<float4> s((size_t)P * PairCfg<P>::LD), tw(2 * P);
        fill_twiddles4<P>(ctx, tw.data());
        fill_row_twiddles4<P>(ctx, tw.data(), tw.data() + P);
        for (int ta = 0; ta < prm.tiles_total; ta += 2) {
            const bool b_valid = ta + 1 < prm.tiles_total;
            const int tb = b_valid ? ta + 1 : ta;
            float2 a = make_float2(0.f, 0.f), p = make_float2(0.f, 0.f);
            pair_process<P, T, LUMA3>(ctx, prm, ta, tb, b_valid, s.data(), tw.data(), a, p);
            prm.partials[2 * ta] = a.x;
            prm.partials[2 * ta + 1] = p.x;
            if (b_valid) {
                prm.partials[2 * tb] = a.y;
                prm.partials[2 * tb + 1] = p.y;
            }
        }
    }
}

template <typename T, bool LUMA3>
void run_sub(Params prm) {
    SerialCtx ctx;
    const int D = prm.sub_d, npp = D * D / 2;
    std::vector<float2> s((size_t)2 * 64 * SubCfg::LD);
    for (int base = 0; base < prm.tiles_total; base += prm.chunk_tiles) {
        prm.tile_base = base;
        prm.chunk_now = prm.tiles_total - base < prm.chunk_tiles ? prm.tiles_total - base : prm.chunk_tiles;
        if (D == 4) {  // the cluster variant's load (sub_fwd_load_quad): both halves, both column pairs, then the transforms
            std::vector<float2> s2((size_t)2 * 64 * SubCfg::LD);
            for (int w = 0; w < prm.chunk_now * 4; ++w) {
                const TileCoord tc = decode_tile(prm, base + (w >> 2));
                for (int half = 0; half < 2; ++half) sub_fwd_load_quad<T, LUMA3>(ctx, prm, tc, w & 3, half, s.data(), s2.data());
                for (int i = 0; i < 2; ++i) {
                    SubUnit su;
                    su.tile_local = w >> 2;
                    su.p = w & 3;
                    su.i = i;
                    su.plane = su.p * 2 + i;
                    float2* t = i ? s2.data() : s.data();
                    sub_fwd_rows(ctx, t);
                    sub_fwd_cols_store(ctx, prm, su, t);
                }
            }
        } else if (D == 8) {  // the 4-CTA cluster variant's load (sub_fwd_load_oct): four quarters, four column pairs
            std::vector<float2> tb[4];
            for (auto& t : tb) t.resize((size_t)2 * 64 * SubCfg::LD);
            float2* const dst[4] = {tb[0].data(), tb[1].data(), tb[2].data(), tb[3].data()};
            for (int w = 0; w < prm.chunk_now * 8; ++w) {
                const TileCoord tc = decode_tile(prm, base + (w >> 3));
                for (int quarter = 0; quarter < 4; ++quarter) sub_fwd_load_oct<T, LUMA3>(ctx, prm, tc, w & 7, quarter, dst);
                for (int i = 0; i < 4; ++i) {
                    SubUnit su;
                    su.tile_local = w >> 3;
                    su.p = w & 7;
                    su.i = i;
                    su.plane = su.p * 4 + i;
                    sub_fwd_rows(ctx, dst[i]);
                    sub_fwd_cols_store(ctx, prm, su, dst[i]);
                }
            }
        } else {
            for (int u = 0; u < prm.chunk_now * npp; ++u) sub_fwd_process<T, LUMA3>(ctx, prm, u, s.data());
        }
        if (D == 8) {
            std::vector<float2> sm(Combine8Cfg::SMEM / sizeof(float2));
            for (int lt = 0; lt < prm.chunk_now; ++lt)
                for (int row = 0; row < kCombine8Parts; ++row) {
                    float a = 0.f, p = 0.f;
                    combine8_rows(ctx, prm, sub_plane(prm, lt, 0), row, sm.data(), a, p);
                    prm.partials[2 * ((size_t)(base + lt) * kCombine8Parts + row)] = a;
                    prm.partials[2 * ((size_t)(base + lt) * kCombine8Parts + row) + 1] = p;
                }
        }
        for (int lt = 0; lt < prm.chunk_now && D != 8; ++lt) {
            float2* ws_tile = sub_plane(prm, lt, 0);
            for (int part = 0; part < kCombineParts; ++part) {
                float a = 0.f, p = 0.f;
                for (int item = part * kCombineItemsPerPart; item < (part + 1) * kCombineItemsPerPart && item < kCombineItems; ++item) {
                    if (D == 2) combine_item<2>(prm, ws_tile, item, a, p);
                    else combine_item<4>(prm, ws_tile, item, a, p);
                }
                prm.partials[2 * ((size_t)(base + lt) * kCombineParts + part)] = a;
                prm.partials[2 * ((size_t)(base + lt) * kCombineParts + part) + 1] = p;
            }
        }
        if (prm.grad && D == 4) {  // the cluster variant's store (sub_inv_store_quad)
            std::vector<float2> t0((size_t)64 * SubCfg::LD), t1((size_t)64 * SubCfg::LD);
            for (int w = 0; w < prm.chunk_now * 4; ++w) {
                for (int i = 0; i < 2; ++i) {
                    SubUnit su;
                    su.tile_local = w >> 2;
                    su.p = w & 3;
                    su.i = i;
                    su.plane = su.p * 2 + i;
                    float2* t = i ? t1.data() : t0.data();
                    sub_inv_cols(ctx, prm, su, t);
                    sub_inv_rows(ctx, t);
                }
                const TileCoord tc = decode_tile(prm, base + (w >> 2));
                // both store forms get checked against the oracle: odd tiles take the product form (one CTA, both planes),
                // even tiles the 2-CTA cluster form ((own tile, peer tile) of the CTA with each rank)
                if ((w >> 2) & 1) {
                    sub_inv_store_rows4<T, LUMA3>(ctx, prm, tc, w & 3, t0.data(), t1.data());
                } else {
                    for (int half = 0; half < 2; ++half)
                        sub_inv_store_quad<T, LUMA3>(ctx, prm, tc, w & 3, half, half == 0 ? t0.data() : t1.data(), half == 0 ? t1.data() : t0.data());
                }
            }
        } else if (prm.grad && D == 8) {  // the 4-CTA cluster variant's store (sub_inv_store_oct)
            std::vector<float2> tb[4];
            for (auto& t : tb) t.resize((size_t)64 * SubCfg::LD);
            const float2* const src[4] = {tb[0].data(), tb[1].data(), tb[2].data(), tb[3].data()};
            for (int w = 0; w < prm.chunk_now * 8; ++w) {
                for (int i = 0; i < 4; ++i) {
                    SubUnit su;
                    su.tile_local = w >> 3;
                    su.p = w & 7;
                    su.i = i;
                    su.plane = su.p * 4 + i;
                    sub_inv_cols(ctx, prm, su, tb[i].data());
                    sub_inv_rows(ctx, tb[i].data());
                }
                const TileCoord tc = decode_tile(prm, base + (w >> 3));
                for (int quarter = 0; quarter < 4; ++quarter) sub_inv_store_oct<T, LUMA3>(ctx, prm, tc, w & 7, quarter, src);
            }
        } else if (prm.grad) {
            for (int u = 0; u < prm.chunk_now * npp; ++u) sub_inv_process<T, LUMA3>(ctx, prm, u, s.data());
        }
    }
}

template <int P, typename T, bool LUMA3>
void run_split(Params prm) {
    if constexpr (P >= 64) {
        using Sp = Split<P>;
        SerialCtx ctx;
        std::vector<float2> s((size_t)P * (2 * Sp::GS + 1) + (size_t)Sp::RS * (P + 1)), tw(P);
        fill_twiddles<P>(ctx, tw.data());
        for (int base = 0; base < prm.tiles_total; base += prm.chunk_tiles) {
            prm.tile_base = base;
            const int nt = prm.tiles_total - base < prm.chunk_tiles ? prm.tiles_total - base : prm.chunk_tiles;
            for (int lt = 0; lt < nt; ++lt)
                for (int sl = 0; sl < Sp::ROW_SLABS; ++sl) split_rows_fwd<P, T, LUMA3>(ctx, prm, lt, sl, s.data(), tw.data());
            for (int lt = 0; lt < nt; ++lt)
                for (int pr = 0; pr < Sp::PARTS; ++pr) {
                    float a = 0.f, p = 0.f;
                    split_cols<P>(ctx, prm, lt, pr, s.data(), tw.data(), a, p);
                    prm.partials[2 * ((size_t)(base + lt) * Sp::PARTS + pr)] = a;
                    prm.partials[2 * ((size_t)(base + lt) * Sp::PARTS + pr) + 1] = p;
                }
            if (prm.grad)
                for (int lt = 0; lt < nt; ++lt)
                    for (int sl = 0; sl < Sp::ROW_SLABS; ++sl) split_rows_inv<P, T, LUMA3>(ctx, prm, lt, sl, s.data(), tw.data());
        }
    }
}

template <int P, typename T, bool LUMA3>
void run(const Params& prm, bool split) {
    if ((P == 128 || P == 256 || P == 512) && prm.sub_d > 1) run_sub<T, LUMA3>(prm);
    else if (split) run_split<P, T, LUMA3>(prm);
    else if (P == 64 && pair_supported(prm) && !(prm.flags & TFCFFT_USE_PAIR)) run_line<T, LUMA3>(prm);
    else if (P == 64 && pair_supported(prm)) run_pair<P, T, LUMA3>(prm);
    else if constexpr (P <= 128) run_resident<P, T, LUMA3>(prm);
}

template <int P, typename T>
void run_l(const Params& prm, bool split, bool luma3) {
    if (luma3) run<P, T, true>(prm, split);
    else run<P, T, false>(prm, split);
}

}  // namespace

void TFC_FN(emu_run)(Params& prm, const Geometry& g) {
    switch (g.p) {
        case 16: run_l<16, TFC_T>(prm, g.split, g.luma3); break;
        case 32: run_l<32, TFC_T>(prm, g.split, g.luma3); break;
        case 64: run_l<64, TFC_T>(prm, g.split, g.luma3); break;
        case 128: run_l<128, TFC_T>(prm, g.split, g.luma3); break;
        case 256: run_l<256, TFC_T>(prm, g.split, g.luma3); break;
        case 512: run_l<512, TFC_T>(prm, g.split, g.luma3); break;
    }
}

