// sub_ring.cuh -- forward launch of the sub-tile pipeline for D = 4 (256 x 256 tiles) on the asynchronous ring engine of
// line_ring.cuh.
//
// Why: `sub_fwd4_kernel` is 47 % of the global-FFT step (35.7 of 79 us, profiles/r02_launches_global256.csv) and it is
// a load phase: four dependent rounds of register-staged loads per unit (`long_scoreboard` 4.2 + `lg_throttle` 4.5 per
// issue), one wave of 256 cluster work items on 222 cluster slots, a 15 us tail.  Here ONE persistent CTA per SM owns a
// whole (tile, row phase p) super-unit = the four sub-images (p, q = 0..3):
//   * a producer warp streams the rows 4a + p of the tile -- full 256-pixel rows, all channels, fake and real -- with
//     TMA tensor copies whose tensor map has a traversal stride of 4 along H (box = 4 strided rows per tensor, 24 KB
//     slabs for fp32 RGB) into a 3-slot ring: HBM latency is covered by the ring, every 32-byte sector is fetched once;
//   * all 256 worker threads fold a slab to luma and scatter the four column phases into the four work tiles;
//   * each 64-thread group then transforms ITS sub-image (rows, columns: one rolled copy of the 64-point core) and
//     writes its plane of the L2 workspace, while the producer already stages the next super-unit.
// Same arithmetic, same order as sub_fwd4_kernel: the planes -- and so the loss and the gradient -- are bit-identical.
#pragma once
#include "line_ring.cuh"
#include "sub_tile.cuh"

namespace tfcfft {

struct SubRingCfg {
    static constexpr int G = 4;                        // one group per column phase q
    static constexpr int RING = 3;
    static constexpr int RPS = 4;                      // strided rows per slab and tensor
    static constexpr int SLABS = 64 / RPS;             // per super-unit
    static constexpr int NT = 64 * G + 32;
    static constexpr int TILE_BYTES = (int)SubCfg::SMEM_INV;  // 64 x 65 float2
    template <typename T, int NC> static constexpr int slab_bytes() { return 2 * NC * RPS * 256 * (int)sizeof(T); }
    static constexpr int SLOT_BYTES = 2 * 3 * RPS * 256 * 4;  // fp32 RGB: 24576
    static constexpr size_t SMEM = (size_t)RING * SLOT_BYTES + (size_t)G * TILE_BYTES + 256;
};

#ifdef __CUDACC__
template <typename T, bool LUMA3>
__global__ void __launch_bounds__(SubRingCfg::NT, 1) sub_fwd_ring_kernel(const __grid_constant__ Params prm,
                                                                         const __grid_constant__ CUtensorMap map_fake,
                                                                         const __grid_constant__ CUtensorMap map_real) {
    using C = SubRingCfg;
    constexpr int NC = LUMA3 ? 3 : 1, LD = SubCfg::LD, RPS = C::RPS, RING = C::RING;
    constexpr int SLAB = C::slab_bytes<T, NC>(), HALF = SLAB / 2, ROWB = 256 * (int)sizeof(T), PX4 = 4 * (int)sizeof(T);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* ring = smem_raw;
    float2* tiles = reinterpret_cast<float2*>(smem_raw + RING * C::SLOT_BYTES);
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem_raw + RING * C::SLOT_BYTES + C::G * C::TILE_BYTES);
    const unsigned full0 = smem_u32(bars), empty0 = smem_u32(bars + RING);
    const int tid = (int)threadIdx.x;
    if (tid == 0) {
        for (int i = 0; i < RING; ++i) {
            mbar_init(full0 + 8 * i, 1);
            mbar_init(empty0 + 8 * i, 64 * C::G);  // every worker thread reads every slab
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    pdl_wait();
    const int nsuper = prm.chunk_now * 4;  // (tile, row phase)
    const int mine = (nsuper - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    if (tid >= 64 * C::G) {
        if (tid == 64 * C::G) {
            const unsigned long long pol = policy_evict_first();
            unsigned c = 0;
            for (int k = 0; k < mine; ++k) {
                const int su = (int)blockIdx.x + k * (int)gridDim.x;
                const TileCoord tc = decode_tile(prm, prm.tile_base + (su >> 2));
                const int p = su & 3;
                for (int i = 0; i < C::SLABS; ++i, ++c) {
                    const unsigned slot = c % RING, ph = (c / RING) & 1;
                    mbar_wait(empty0 + 8 * slot, ph ^ 1);
                    const unsigned fb = full0 + 8 * slot;
                    mbar_arrive_expect_tx(fb, SLAB);
                    const unsigned dst = smem_u32(ring + slot * C::SLOT_BYTES);
                    const int x = tc.px * 256, y = tc.py * 256 + 4 * (RPS * i) + p;  // rows y, y + 4, y + 8, y + 12
                    tma_load_4d(dst, &map_fake, x, y, tc.ch, tc.n, fb, pol);           // [channel][row][256 px]
                    tma_load_4d(dst + HALF, &map_real, x, y, tc.ch, tc.n, fb, pol);
                }
            }
        }
    } else {
        const int q = tid >> 6, gtid = tid & 63;
        const RingWorkerCtx gctx{gtid, q + 1};
        float2* mytile = tiles + (size_t)q * (C::TILE_BYTES / sizeof(float2));
        const bool quant = (prm.flags & TFCFFT_QUANTIZE_U8) != 0;
        unsigned c = 0;
        for (int k = 0; k < mine; ++k) {
            const int su = (int)blockIdx.x + k * (int)gridDim.x;
            SubUnit unit;
            unit.tile_local = su >> 2;
            unit.p = su & 3;
            for (int i = 0; i < C::SLABS; ++i, ++c) {
                const unsigned slot = c % RING, ph = (c / RING) & 1;
                mbar_wait(full0 + 8 * slot, ph);
                const unsigned char* slab = ring + slot * C::SLOT_BYTES;
                // one 4-pixel item per thread: slab row r (sub-image row a = RPS * i + r), pixels 4 b .. 4 b + 3
                const int b = tid & 63, r = tid >> 6, a = RPS * i + r;
                float raw[2][NC][4];
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int ch = 0; ch < NC; ++ch) SlabIO<T>::load4(slab + h * HALF + (ch * RPS + r) * ROWB + b * PX4, raw[h][ch]);
                mbar_arrive(empty0 + 8 * slot);  // the slab is in registers
                float z[2][4];
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int qq = 0; qq < 4; ++qq) {
                        if (!quant) {
                            float f = prm.lw[0] * raw[h][0][qq];
                            if constexpr (LUMA3) f = fmaf(prm.lw[2], raw[h][2][qq], fmaf(prm.lw[1], raw[h][1][qq], f));
                            z[h][qq] = f;
                        } else if constexpr (LUMA3) {
                            z[h][qq] = (float)((19595 * IO<T>::quant(raw[h][0][qq]) + 38470 * IO<T>::quant(raw[h][1][qq]) +
                                                7471 * IO<T>::quant(raw[h][2][qq]) + 0x8000) >> 16);
                        } else {
                            z[h][qq] = (float)IO<T>::quant(raw[h][0][qq]);
                        }
                    }
#pragma unroll
                for (int qq = 0; qq < 4; ++qq)
                    tiles[(size_t)qq * (C::TILE_BYTES / sizeof(float2)) + a * LD + b] = make_float2(z[0][qq], z[1][qq]);
            }
            bar_sync(6, 64 * C::G);  // all four work tiles are complete
            float2* plane = sub_plane(prm, unit.tile_local, unit.p * 4 + q);
#pragma unroll 1
            for (int pass = 0; pass < 2; ++pass) {  // rolled: one copy of the 64-point core
                float2 v[64];
                float2* row = mytile + gtid * LD;
                if (pass == 0) {
#pragma unroll
                    for (int x = 0; x < 64; ++x) v[x] = row[x];
                } else {
#pragma unroll
                    for (int y = 0; y < 64; ++y) v[y] = mytile[y * LD + gtid];
                }
                fft64<false>(v);
                if (pass == 0) {
#pragma unroll
                    for (int sl = 0; sl < 64; ++sl) row[fft64_freq(sl)] = v[sl];
                    gctx.sync();
                } else {
#pragma unroll
                    for (int sl = 0; sl < 64; ++sl) plane[fft64_freq(sl) * 64 + gtid] = v[sl];
                }
            }
            bar_sync(6, 64 * C::G);  // the next super-unit's pixels overwrite all four tiles
        }
    }
    pdl_release();
}

// returns TFCFFT_ERR_STRIDE when the inputs cannot be described by a tensor map
template <typename T, bool LUMA3>
int launch_sub_fwd_ring(const Params& prm, cudaStream_t st) {
    constexpr int NC = LUMA3 ? 3 : 1;
    auto kernel = sub_fwd_ring_kernel<T, LUMA3>;
    static KernelFacts facts;
    if (int rc = facts.get(kernel, SubRingCfg::NT, SubRingCfg::SMEM, nullptr)) return rc;
    alignas(64) CUtensorMap mf, mr;
    if (!make_tile_map<T>(&mf, prm.fake, prm.fs, prm, NC, SubRingCfg::RPS, 256, 4) ||
        !make_tile_map<T>(&mr, prm.real, prm.rs, prm, NC, SubRingCfg::RPS, 256, 4))
        return TFCFFT_ERR_STRIDE;
    const int sms = device_sms(), nsuper = prm.chunk_now * 4;
    const int grid = nsuper < sms ? nsuper : sms;
    static const bool off = getenv("TFCFFT_NO_PDL") != nullptr;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)SubRingCfg::NT);
    cfg.dynamicSmemBytes = SubRingCfg::SMEM;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = off ? 0 : 1;
    if (cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, prm, mf, mr)) return (int)e;
    g_launches++;
    return 0;
}
#endif  // __CUDACC__

}  // namespace tfcfft
