// line_ring.cuh -- 64 x 64 tiles: the thread-per-line transforms of line_tile.cuh fed by an ASYNCHRONOUS shared-memory
// ring (sm_90+ bulk copies, mbarrier completion) instead of blocking per-thread loads.
//
// Why (profiles/r01_ncu_patch16_line_kernel_summary.txt): in line_kernel every CTA runs load -> transforms -> store
// in sequence and the bytes it can keep in flight are bounded by its registers (24 KB per 64-thread CTA, four
// dependent HBM round trips per tile); the memory phases and the transforms did not overlap (long_scoreboard was the
// top stall, DRAM 38 % busy, issue slots 39 % busy).  Here ONE persistent CTA per SM holds
//   * a producer warp whose elected lane streams raw row slabs of the tiles (fake + real, all channels) into a ring
//     of RING slots with TMA tensor copies (`cp.async.bulk.tensor.4d...mbarrier::complete_tx::bytes` over a rank-4
//     tensor map of the NCHW input, box [1, C, rows, 64]; SASS: UTMALDG) -- the bytes in flight are bounded by the
//     ring (RING x 12 KB), not by registers, and never stall a compute warp.  (A first version issued one 256-byte
//     `cp.async.bulk` per pixel row, 384 per tile: the copy issue rate, ~65 cycles each, capped the kernel at
//     0.71 M img/s whatever the ring shape -- profiles/r02_ring_ab.txt.)
//   * G worker groups of 64 threads (one tile each, round robin): wait on a slot's `full` mbarrier, fold the slab to
//     luma (`z = fake + i real`) into their work tile, release the slot (`empty` mbarrier), then run the register-
//     resident 64-point line transforms, the loss stage and the inverse transforms of line_tile.cuh unchanged, and
//     write the gradient tile with coalesced 128-bit streaming stores.
// The slabs of consecutive tiles go to consecutive workers, so the workers' load phases are staggered by
// construction and the transforms of G - 1 tiles overlap the HBM stream of the next one.
#pragma once
#include <cuda.h>

#include <string>

#include "kernel_common.cuh"
#include "line_tile.cuh"

namespace tfcfft {

#ifndef TFCFFT_RING_WORKERS
#define TFCFFT_RING_WORKERS 4
#endif
#ifndef TFCFFT_RING_SLOTS
#define TFCFFT_RING_SLOTS 7
#endif
// (an L2 look-ahead with TMA prefetches, 1-2 tiles per CTA, was measured and lost 1-4 %: the ring already covers HBM latency)

constexpr int kRingSlotBytes = 12288;  // one slab: fake + real, NC channels, RPS rows of 64 pixels
template <int G_, int RING_, int NTG_ = 64>
struct RingCfgT {
    static constexpr int G = G_;                       // worker groups (tiles in flight) per CTA
    static constexpr int RING = RING_;                 // raw slabs in flight
    static constexpr int NTG = NTG_;                   // threads per worker group: 64 = thread per line, 128 = half-line engine
    static constexpr int SLOT_BYTES = kRingSlotBytes;
    static constexpr int NT = NTG * G + 32;            // workers + producer warp
    static constexpr int TILE_BYTES = (int)LineCfg::SMEM;  // 64 x 65 float2
    static constexpr int BAR_BYTES = 256;              // full[RING], empty[RING] (8 bytes each)
    static constexpr size_t SMEM = (size_t)G * TILE_BYTES + (size_t)RING * SLOT_BYTES + BAR_BYTES + 32 * G;
    static_assert(2 * RING * 8 + 8 <= BAR_BYTES, "barrier area too small");
    static_assert(SMEM + 2048 <= 227 * 1024, "ring configuration exceeds the shared memory of an SM");
};
using RingCfg = RingCfgT<TFCFFT_RING_WORKERS, TFCFFT_RING_SLOTS>;  // default configuration

// rows per slab so that one slab fits a slot: 2 tensors x NC channels x RPS rows x 64 pixels
template <typename T, int NC>
struct RingSlab {
    static constexpr int ROW_BYTES = 64 * (int)sizeof(T);
    static constexpr int RPS_RAW = kRingSlotBytes / (2 * NC * ROW_BYTES);
    static constexpr int RPS = RPS_RAW >= 64 ? 64 : RPS_RAW >= 32 ? 32 : RPS_RAW >= 16 ? 16 : 8;
    static constexpr int SLABS = 64 / RPS;              // slabs per tile
    static constexpr int COPIES = 2 * NC * RPS;         // row copies per slab
    static constexpr int BYTES = COPIES * ROW_BYTES;
    static_assert(BYTES <= kRingSlotBytes, "slab does not fit its slot");
};

// The bulk-copy engine needs 16-byte aligned global rows: base pointers and all strides in bytes multiples of 16
// (always true for fp32 tensors that passed the C-ABI checks; fp16 / bf16 / uint8 views may miss it).
template <typename T>
inline bool ring_addressable(const Params& prm) {
    auto ok = [](const void* p, const long long* st) {
        if (reinterpret_cast<uintptr_t>(p) % 16) return false;
        for (int i = 0; i < 3; ++i)
            if ((st[i] * (long long)sizeof(T)) % 16) return false;
        return true;
    };
    return prm.real_q[0] == nullptr && ok(prm.fake, prm.fs) && ok(prm.real, prm.rs);
}

#ifdef __CUDACC__
// ---- mbarrier / bulk-copy primitives (PTX ISA 8.x, sm_90+) -------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(bar),
        "r"(parity)
        : "memory");
}
// TMA tensor copy global -> shared: box of the rank-4 map at (x, y, c, n), completion counted in bytes on `bar`;
// L2 evict-first (the pixels are read once)
__device__ __forceinline__ void tma_load_4d(unsigned dst, const CUtensorMap* map, int x, int y, int c, int n, unsigned bar,
                                            unsigned long long policy) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4, %5}], [%6], %7;" ::"r"(dst),
        "l"(reinterpret_cast<unsigned long long>(map)), "r"(x), "r"(y), "r"(c), "r"(n), "r"(bar), "l"(policy)
        : "memory");
}
__device__ __forceinline__ unsigned long long policy_evict_first() {
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}

// plain (non-streaming) 4-pixel read of a slab row in shared memory
template <typename T> struct SlabIO;
template <> struct SlabIO<float> {
    __device__ __forceinline__ static void load4(const unsigned char* p, float* v) {
        const float4 t = *reinterpret_cast<const float4*>(p);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
};
template <> struct SlabIO<__half> {
    __device__ __forceinline__ static void load4(const unsigned char* p, float* v) {
        const uint2 t = *reinterpret_cast<const uint2*>(p);
        const __half2 a = *reinterpret_cast<const __half2*>(&t.x), b = *reinterpret_cast<const __half2*>(&t.y);
        v[0] = __low2float(a); v[1] = __high2float(a); v[2] = __low2float(b); v[3] = __high2float(b);
    }
};
template <> struct SlabIO<__nv_bfloat16> {
    __device__ __forceinline__ static void load4(const unsigned char* p, float* v) {
        const uint2 t = *reinterpret_cast<const uint2*>(p);
        v[0] = __uint_as_float(t.x << 16); v[1] = __uint_as_float(t.x & 0xFFFF0000u);
        v[2] = __uint_as_float(t.y << 16); v[3] = __uint_as_float(t.y & 0xFFFF0000u);
    }
};
template <> struct SlabIO<uint8_t> {
    __device__ __forceinline__ static void load4(const unsigned char* p, float* v) {
        const uchar4 t = *reinterpret_cast<const uchar4*>(p);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
};

// execution context of one 64-thread worker group: its own named barrier
template <int NTG>
struct RingWorkerCtxT {
    int tid, bar;
    static constexpr int nthreads = NTG;
    __device__ __forceinline__ void sync() const { bar_sync(bar, NTG); }
    __device__ __forceinline__ void warp_sync() const { __syncwarp(); }
    __device__ __forceinline__ void mark(int) const {}
};
using RingWorkerCtx = RingWorkerCtxT<64>;

// ---- worker: one slab (rows y0 .. y0 + RPS - 1 of the tile) -> luma -> work tile ---------------------------------
// slab layout: [fake | real][channel][row][64 pixels]; work tile layout as line_load (line_slot permutation).
// Returns false as soon as a fake pixel differs from the real pixel (after luma): the tile-level `fake == real` test.
template <typename T, bool LUMA3, int NTG = 64>
__device__ __forceinline__ bool ring_convert(const Params& prm, const unsigned char* slab, int y0, int tid, float2* s) {
    constexpr int NC = LUMA3 ? 3 : 1, LD = LineCfg::LD;
    using SL = RingSlab<T, NC>;
    constexpr int RPS = SL::RPS, RB = SL::ROW_BYTES, PX4 = 4 * (int)sizeof(T);
    const bool quant = (prm.flags & TFCFFT_QUANTIZE_U8) != 0;
    bool same = true;
    constexpr int ITEMS = RPS * 16, PER = ITEMS / NTG;  // 4-pixel items per thread
    static_assert(ITEMS % NTG == 0, "slab items must divide over the group");
#pragma unroll
    for (int u = 0; u < PER; ++u) {
        const int it = tid + NTG * u, x4 = it & 15, r = it >> 4;
        float raw[2][NC][4];
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int c = 0; c < NC; ++c) SlabIO<T>::load4(slab + ((h * NC + c) * RPS + r) * RB + x4 * PX4, raw[h][c]);
        float2* row = s + (y0 + r) * LD + x4;
        if (!quant) {
            // luma of two pixels per packed instruction: (w0 * r + w1 * g + w2 * b) on (pixel i, pixel i + 1) lanes
#pragma unroll
            for (int i = 0; i < 4; i += 2) {
                float2 z[2];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    float2 f = p_mul(p_dup(prm.lw[0]), make_float2(raw[h][0][i], raw[h][0][i + 1]));
                    if constexpr (LUMA3) {
                        f = p_fma(p_dup(prm.lw[1]), make_float2(raw[h][1][i], raw[h][1][i + 1]), f);
                        f = p_fma(p_dup(prm.lw[2]), make_float2(raw[h][2][i], raw[h][2][i + 1]), f);
                    }
                    z[h] = f;
                }
                same = same && (z[0].x == z[1].x) && (z[0].y == z[1].y);
                row[16 * i] = make_float2(z[0].x, z[1].x);  // == line_slot(4 * x4 + i)
                row[16 * (i + 1)] = make_float2(z[0].y, z[1].y);
            }
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float z[2];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    if constexpr (LUMA3) {
                        z[h] = (float)((19595 * IO<T>::quant(raw[h][0][i]) + 38470 * IO<T>::quant(raw[h][1][i]) +
                                        7471 * IO<T>::quant(raw[h][2][i]) + 0x8000) >> 16);
                    } else {
                        z[h] = (float)IO<T>::quant(raw[h][0][i]);
                    }
                }
                same = same && (z[0] == z[1]);
                row[16 * i] = make_float2(z[0], z[1]);
            }
        }
    }
    return same;
}

// Gradient store of an all-zero tile (fake == real): the loss terms and their gradient vanish identically.
template <typename T, bool LUMA3, class Ctx>
__device__ __forceinline__ void ring_store_zero(const Ctx& ctx, const Params& prm, const TileCoord& tc) {
    constexpr int NC = LUMA3 ? 3 : 1;
    T* gp = const_cast<T*>(tile_ptr<T>(prm.grad, prm.gs, tc, 64));
    const int sh = (int)prm.gs[2], sc = (int)prm.gs[1];
    if (prm.flags & TFCFFT_GRAD_ACCUMULATE) return;  // adding zero
    const float z[4] = {0.f, 0.f, 0.f, 0.f};
    for (int it = ctx.tid; it < 64 * 16; it += ctx.nthreads) {
        const int x = (it & 15) * 4, y = it >> 4;
#pragma unroll
        for (int c = 0; c < NC; ++c) IO<T>::store4(gp + y * sh + c * sc + x, z);
    }
}

template <typename T, bool LUMA3, class RingCfg>
__global__ void __launch_bounds__(RingCfg::NT, 1) line_ring_kernel(const __grid_constant__ Params prm,
                                                                    const __grid_constant__ CUtensorMap map_fake,
                                                                    const __grid_constant__ CUtensorMap map_real) {
    constexpr int G = RingCfg::G, RING = RingCfg::RING, NC = LUMA3 ? 3 : 1, NTG = RingCfg::NTG;
    using SL = RingSlab<T, NC>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* ring = smem_raw;                                                   // RING slots
    float2* tiles = reinterpret_cast<float2*>(smem_raw + RING * RingCfg::SLOT_BYTES);  // G work tiles
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem_raw + RING * RingCfg::SLOT_BYTES + G * RingCfg::TILE_BYTES);
    float* scratch = reinterpret_cast<float*>(bars + RingCfg::BAR_BYTES / 8);          // [G][4]: cross-warp sums
    // tile of this CTA whose worker may take slabs off the ring.  The mbarrier waits below only carry ONE bit of
    // phase, so a worker must not start waiting for its first slab while the ring is still several rounds behind
    // (it would wake up on an earlier round of the same slot): workers take turns in tile order.
    volatile unsigned* turn = reinterpret_cast<volatile unsigned*>(bars + 2 * RING);
    const unsigned full0 = smem_u32(bars), empty0 = smem_u32(bars + RING);
    const int tid = (int)threadIdx.x;
    if (tid == 0) {
        for (int i = 0; i < RING; ++i) {
            mbar_init(full0 + 8 * i, 1);    // one arrive.expect_tx by the producer + the copies' bytes
            mbar_init(empty0 + 8 * i, NTG);  // every thread of the consuming group
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        *turn = 0u;
    }
    __syncthreads();
    pdl_wait();
    // tiles of this CTA: blockIdx.x + k * gridDim.x, k = 0 .. ntiles - 1; tile k belongs to worker k % G
    const int ntiles = ((int)prm.tiles_total - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    if (tid >= NTG * G) {
        // ---------------- producer warp: raw row slabs -> ring (one elected lane issues the TMA copies) --------
        if (tid == NTG * G) {
            const unsigned long long pol = policy_evict_first();
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<unsigned long long>(&map_fake)) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<unsigned long long>(&map_real)) : "memory");
            unsigned c = 0;  // slab counter
            for (int k = 0; k < ntiles; ++k) {
                const TileCoord tc = decode_tile(prm, (int)blockIdx.x + k * (int)gridDim.x);
                for (int i = 0; i < SL::SLABS; ++i, ++c) {
                    const unsigned slot = c % RING, ph = (c / RING) & 1;
                    mbar_wait(empty0 + 8 * slot, ph ^ 1);  // the consumers of the previous round have released the slot
                    const unsigned fb = full0 + 8 * slot;
                    mbar_arrive_expect_tx(fb, SL::BYTES);
                    const unsigned dst0 = smem_u32(ring + slot * RingCfg::SLOT_BYTES);
                    const int x = tc.px * 64, y = tc.py * 64 + i * SL::RPS;
                    tma_load_4d(dst0, &map_fake, x, y, tc.ch, tc.n, fb, pol);                   // [channel][row][64 px]
                    tma_load_4d(dst0 + SL::BYTES / 2, &map_real, x, y, tc.ch, tc.n, fb, pol);
                }
            }
        }
    } else {
        // ---------------- worker groups ----------------
        const int w = tid / NTG, gtid = tid % NTG;
        const RingWorkerCtxT<NTG> ctx{gtid, w + 1};  // named barrier id w + 1 (id 0 is __syncthreads)
        float2* s = tiles + (size_t)w * (RingCfg::TILE_BYTES / sizeof(float2));
        float* red = scratch + 8 * w;
        const bool want_grad = prm.grad != nullptr;
        for (int k = w; k < ntiles; k += G) {
            const int tile = (int)blockIdx.x + k * (int)gridDim.x;
            const TileCoord tc = decode_tile(prm, tile);
            bool same = true;
            unsigned c = (unsigned)k * SL::SLABS;
            if (gtid == 0) {
                while (*turn != (unsigned)k) __nanosleep(64);
            }
            ctx.sync();  // my turn: every slab before tile k's has been taken
            for (int i = 0; i < SL::SLABS; ++i, ++c) {
                const unsigned slot = c % RING, ph = (c / RING) & 1;
                mbar_wait(full0 + 8 * slot, ph);
                if (i == SL::SLABS - 1 && gtid == 0) *turn = (unsigned)k + 1u;  // the next worker may start waiting
                same = ring_convert<T, LUMA3, NTG>(prm, ring + slot * RingCfg::SLOT_BYTES, i * SL::RPS, gtid, s) && same;
                mbar_arrive(empty0 + 8 * slot);
            }
            // group vote: is fake == real on the whole tile?  (the barrier also publishes the work tile)
            int all_same;
            asm volatile(
                "{\n\t.reg .pred p, q;\n\tsetp.ne.s32 p, %1, 0;\n\tbar.red.and.pred q, %2, %3, p;\n\tselp.s32 %0, 1, 0, q;\n\t}"
                : "=r"(all_same)
                : "r"((int)same), "r"(w + 1), "r"(NTG)
                : "memory");
            float a = 0.f, p = 0.f;
            if (all_same) {
                // identical inputs: every loss term is exactly zero and so is the gradient (reference semantics of
                // L1Loss / sign(0) = 0), independent of rounding in the packed transform
                if (want_grad) ring_store_zero<T, LUMA3>(ctx, prm, tc);
                ctx.sync();
            } else {
                const int npass = want_grad ? 4 : 2;
#pragma unroll 1
                for (int pass = 0; pass < npass; ++pass) {  // rolled: ONE copy of the transform core
                    if constexpr (NTG == 128) line2_fft_pass(ctx, s, pass);  // two threads per line, 32-point core
                    else line_fft_pass(ctx, s, pass);
                    ctx.sync();
                    if (pass == 1) {
                        line_bins(ctx, prm, s, a, p);
                        ctx.sync();
                    }
                }
                if (want_grad) {
                    line_store<T, LUMA3>(ctx, prm, tc, s);
                    ctx.sync();  // the next tile's slabs overwrite the work tile
                }
            }
            // per-tile partial sums: warp shuffle, then the group's two warps through shared memory
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                a += __shfl_down_sync(0xffffffffu, a, o);
                p += __shfl_down_sync(0xffffffffu, p, o);
            }
            if ((gtid & 31) == 0) {
                red[2 * (gtid >> 5)] = a;
                red[2 * (gtid >> 5) + 1] = p;
            }
            ctx.sync();
            if (gtid == 0) {
                float sa = 0.f, sp = 0.f;
#pragma unroll
                for (int i = 0; i < NTG / 32; ++i) {
                    sa += red[2 * i];
                    sp += red[2 * i + 1];
                }
                prm.partials[2 * tile] = sa;
                prm.partials[2 * tile + 1] = sp;
            }
            ctx.sync();  // `red` is reused by the next tile
        }
    }
    pdl_release();
    finish(prm, gridDim.x);
}

// ---- host: rank-4 tensor maps of the NCHW inputs (driver entry point fetched through the runtime: no libcuda link) --
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeTiledFn tensor_map_encoder() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}
template <typename T> struct TmaType;
template <> struct TmaType<float> { static constexpr CUtensorMapDataType v = CU_TENSOR_MAP_DATA_TYPE_FLOAT32; };
template <> struct TmaType<__half> { static constexpr CUtensorMapDataType v = CU_TENSOR_MAP_DATA_TYPE_FLOAT16; };
template <> struct TmaType<__nv_bfloat16> { static constexpr CUtensorMapDataType v = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16; };
template <> struct TmaType<uint8_t> { static constexpr CUtensorMapDataType v = CU_TENSOR_MAP_DATA_TYPE_UINT8; };

// box [1][NC][RPS][64] over [N][C][H][W] with the tensor's own strides (views are fine: strides are multiples of 16 bytes)
template <typename T>
inline bool make_tile_map(CUtensorMap* map, const void* base, const long long* st, const Params& prm, int nc, int rps,
                          int box_w = 64, int row_step = 1) {
    EncodeTiledFn enc = tensor_map_encoder();
    if (!enc) return false;
    const cuuint64_t es = sizeof(T);
    const cuuint64_t dims[4] = {(cuuint64_t)prm.w, (cuuint64_t)prm.h, (cuuint64_t)prm.c, (cuuint64_t)prm.n};
    auto stride = [&](long long s, cuuint64_t lower) {  // a size-1 dimension may carry any stride: give the encoder a valid one
        cuuint64_t b = (cuuint64_t)s * es;
        return b < lower ? lower : b;
    };
    const cuuint64_t row = stride(st[2], dims[0] * es);
    const cuuint64_t chan = stride(st[1], row);
    const cuuint64_t img = stride(st[0], chan);
    const cuuint64_t strides[3] = {row, chan, img};
    // a traversal stride s along H loads every s-th row: the box then spans rps * s rows of the tensor
    const cuuint32_t box[4] = {(cuuint32_t)box_w, (cuuint32_t)(rps * row_step), (cuuint32_t)nc, 1u};
    const cuuint32_t estr[4] = {1u, (cuuint32_t)row_step, 1u, 1u};
    return enc(map, TmaType<T>::v, 4, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// returns TFCFFT_ERR_STRIDE when the inputs cannot be described by a tensor map (the caller falls back to line_kernel)
template <typename T, bool LUMA3, class RingCfg>
int launch_line_ring_cfg(const Params& prm, cudaStream_t st) {
    constexpr int NC = LUMA3 ? 3 : 1;
    using SL = RingSlab<T, NC>;
    auto kernel = line_ring_kernel<T, LUMA3, RingCfg>;
    static KernelFacts facts;
    if (int rc = facts.get(kernel, RingCfg::NT, RingCfg::SMEM, nullptr)) return rc;
    alignas(64) CUtensorMap mf, mr;
    if (!make_tile_map<T>(&mf, prm.fake, prm.fs, prm, NC, SL::RPS) || !make_tile_map<T>(&mr, prm.real, prm.rs, prm, NC, SL::RPS))
        return TFCFFT_ERR_STRIDE;
    const int sms = device_sms();
    // one persistent CTA per SM; with few tiles use fewer CTAs so that every CTA keeps its G workers busy
    long long want = ((long long)prm.tiles_total + RingCfg::G - 1) / RingCfg::G;
    if (want < 1) want = 1;
    const int grid = (int)(want < sms ? want : sms);
    {
        static const bool off = getenv("TFCFFT_NO_PDL") != nullptr;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)grid);
        cfg.blockDim = dim3((unsigned)RingCfg::NT);
        cfg.dynamicSmemBytes = RingCfg::SMEM;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = off ? 0 : 1;
        if (cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, prm, mf, mr)) return (int)e;
    }
    g_launches++;
    return 0;
}

// TFCFFT_RING=<workers>x<slots> selects one of the compiled ring configurations (A/B runs; fp32 objects only)
template <typename T, bool LUMA3>
int launch_line_ring(const Params& prm, cudaStream_t st) {
#if defined(TFC_DT) && TFC_DT == 0
    static const char* sel = getenv("TFCFFT_RING");
    if (sel != nullptr) {
        const std::string v(sel);
        if (v == "4x7") return launch_line_ring_cfg<T, LUMA3, RingCfgT<4, 7>>(prm, st);
        if (v == "5x3") return launch_line_ring_cfg<T, LUMA3, RingCfgT<5, 3>>(prm, st);
        if (v == "5x4") return launch_line_ring_cfg<T, LUMA3, RingCfgT<5, 4>>(prm, st);
        if (v == "4x6") return launch_line_ring_cfg<T, LUMA3, RingCfgT<4, 6>>(prm, st);
        if (v == "4x7") return launch_line_ring_cfg<T, LUMA3, RingCfgT<4, 7>>(prm, st);
        if (v == "h4x7") return launch_line_ring_cfg<T, LUMA3, RingCfgT<4, 7, 128>>(prm, st);
        if (v == "h3x8") return launch_line_ring_cfg<T, LUMA3, RingCfgT<3, 8, 128>>(prm, st);
        if (v == "h4x6") return launch_line_ring_cfg<T, LUMA3, RingCfgT<4, 6, 128>>(prm, st);
    }
#endif
    if (prm.flags & TFCFFT_USE_HALFLINE) return launch_line_ring_cfg<T, LUMA3, RingCfgT<4, 7, 128>>(prm, st);
    // single-channel tiles (per-channel rgb mode, grey inputs): three times the transforms per byte, so the ring can be
    // shallower and a fifth worker pays (measured 0.820 vs 0.786 M img/s on patch-16 rgb b256)
    if constexpr (!LUMA3) return launch_line_ring_cfg<T, LUMA3, RingCfgT<5, 4>>(prm, st);
    else return launch_line_ring_cfg<T, LUMA3, RingCfg>(prm, st);
}
#endif  // __CUDACC__

}  // namespace tfcfft
