// spectral_kernels.cuh -- sm_100a kernels of the generic (all sizes, all modes) path.
//
//   resident_kernel      P <= 128: persistent CTAs, one complex tile resident in shared memory:
//                        load+luma+pack -> row FFT -> column FFT -> loss + spectral gradient ->
//                        inverse column FFT -> inverse row FFT -> gradient store.
//   split_rows_fwd / split_cols / split_rows_inv
//                        P >= 256 (tile does not fit one CTA's shared memory): the row and column
//                        transforms run in separate launches that exchange the complex spectrum
//                        through an L2-sized workspace chunk.
//   grad_scale_kernel    dst = src * scale (autograd's multiplication by grad_output).
//
// Roofline / algorithmic bytes are documented in DESIGN.md.
#pragma once
#include <cooperative_groups.h>

#include "pair_tile.cuh"
#include "sub_tile.cuh"
#include "line_tile.cuh"
#include "triplet.cuh"
#include "regional.cuh"
#include "spectral_core.cuh"

namespace tfcfft {

template <int P> struct ResidentCfg {
    static constexpr int NT = P <= 16 ? 64 : P <= 32 ? 128 : P <= 64 ? 256 : 512;
    static constexpr size_t SMEM = ((size_t)P * (P + 1) + P) * sizeof(float2);
};
template <int P> struct SplitCfg {
    static constexpr int NT = 256;
    static constexpr size_t SMEM_ROWS = ((size_t)Split<P>::RS * (P + 1) + P) * sizeof(float2);
    static constexpr size_t SMEM_COLS = ((size_t)P * (2 * Split<P>::GS + 1) + P) * sizeof(float2);
};

// Block sum of two floats; result valid in thread 0.  Fixed shape -> deterministic.
__device__ __forceinline__ void block_sum2(float& a, float& b) {
    __shared__ float red[2][32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_down_sync(0xffffffffu, a, o);
        b += __shfl_down_sync(0xffffffffu, b, o);
    }
    if (lane == 0) {
        red[0][wid] = a;
        red[1][wid] = b;
    }
    __syncthreads();
    if (wid == 0) {
        a = lane < nw ? red[0][lane] : 0.f;
        b = lane < nw ? red[1][lane] : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a += __shfl_down_sync(0xffffffffu, a, o);
            b += __shfl_down_sync(0xffffffffu, b, o);
        }
    }
    __syncthreads();
}

// Last block to arrive sums all partials in a fixed order (double) and writes the outputs;
// the ticket counter is left at zero for the next call.
__device__ __forceinline__ void finish(const Params& prm, unsigned total_blocks) {
    __shared__ bool last;
    __shared__ double dred[2][32];
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        last = (atomicAdd(prm.counter, 1u) == total_blocks - 1);
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    double a = 0.0, p = 0.0;
    const double img_norm = prm.norm * (double)prm.n;
    for (int img = threadIdx.x; img < prm.n; img += blockDim.x) {
        const float* q = prm.partials + (long long)img * prm.tiles_per_image * prm.parts * 2;
        double ia = 0.0, ip = 0.0;
        for (int i = 0; i < prm.tiles_per_image * prm.parts; ++i) {
            ia += (double)__ldcg(q + 2 * i);
            ip += (double)__ldcg(q + 2 * i + 1);
        }
        if (prm.per_image) {
            prm.per_image[2 * img] = (float)(ia * img_norm);
            prm.per_image[2 * img + 1] = (float)(ip * img_norm);
        }
        a += ia;
        p += ip;
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_down_sync(0xffffffffu, a, o);
        p += __shfl_down_sync(0xffffffffu, p, o);
    }
    if (lane == 0) {
        dred[0][wid] = a;
        dred[1][wid] = p;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        a = 0.0;
        p = 0.0;
        for (int w = 0; w < nw; ++w) {
            a += dred[0][w];
            p += dred[1][w];
        }
        write_outputs(prm, a, p);
        *prm.counter = 0u;
    }
}

template <int P, typename T, bool LUMA3>
__global__ void __launch_bounds__(ResidentCfg<P>::NT) resident_kernel(const __grid_constant__ Params prm) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* s = reinterpret_cast<float2*>(smem_raw);
    float2* tw = s + P * (P + 1);
    const BlockCtx ctx{(int)threadIdx.x, (int)blockDim.x};
    fill_twiddles<P>(ctx, tw);
    ctx.sync();
    for (int tile = blockIdx.x; tile < prm.tiles_total; tile += gridDim.x) {
        float a = 0.f, p = 0.f;
        tile_process<P, T, LUMA3>(ctx, prm, tile, s, tw, a, p);
        block_sum2(a, p);
        if (threadIdx.x == 0) {
            prm.partials[2 * tile] = a;
            prm.partials[2 * tile + 1] = p;
        }
    }
    finish(prm, gridDim.x);
}

// ---- named barriers (bar.sync / bar.arrive with explicit participant counts) ---------------------
__device__ __forceinline__ void bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void bar_arrive(int id, int nthreads) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Execution context of one warp-role group: compile-time size, its own named barrier.
template <int NT, int BAR>
struct GroupCtx {
    int tid;
    static constexpr int nthreads = NT;
    long long* trace;
    __device__ __forceinline__ void sync() const { bar_sync(BAR, NT); }
    __device__ __forceinline__ void warp_sync() const { __syncwarp(); }
    __device__ __forceinline__ void mark(int k) const {
        if (trace != nullptr && tid == 0) trace[k] = clock64();
    }
};

// Sum of four floats over a group of NT threads (NT/32 <= 32 warps); result valid in group thread 0.
template <int NT, int BAR>
__device__ __forceinline__ void group_sum4(int tid, float* red /* [4][32] */, float& a, float& b, float& c, float& d) {
    const int lane = tid & 31, wid = tid >> 5;
    constexpr int NW = NT / 32;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_down_sync(0xffffffffu, a, o);
        b += __shfl_down_sync(0xffffffffu, b, o);
        c += __shfl_down_sync(0xffffffffu, c, o);
        d += __shfl_down_sync(0xffffffffu, d, o);
    }
    if (lane == 0) {
        red[0 * 32 + wid] = a;
        red[1 * 32 + wid] = b;
        red[2 * 32 + wid] = c;
        red[3 * 32 + wid] = d;
    }
    bar_sync(BAR, NT);
    if (wid == 0) {
        a = lane < NW ? red[0 * 32 + lane] : 0.f;
        b = lane < NW ? red[1 * 32 + lane] : 0.f;
        c = lane < NW ? red[2 * 32 + lane] : 0.f;
        d = lane < NW ? red[3 * 32 + lane] : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a += __shfl_down_sync(0xffffffffu, a, o);
            b += __shfl_down_sync(0xffffffffu, b, o);
            c += __shfl_down_sync(0xffffffffu, c, o);
            d += __shfl_down_sync(0xffffffffu, d, o);
        }
    }
    bar_sync(BAR, NT);
}

// Packed tile-pair kernel (pair_tile.cuh), warp-specialised and persistent: one CTA per SM.
//   warps 0..15  (compute): transform the pair resident in work buffer b = i & 1 -- row/column FFTs, loss,
//                           spectral gradient, inverse FFTs, gradient store;
//   warps 16..23 (load):    stream pair i+1 from HBM (coalesced 128-bit loads), fold luma, pack (A, B) and
//                           fill the other work buffer, so HBM latency and the transforms overlap.
//                           and write the finished gradient tiles of pair i-1 out with 128-bit stores.
// Hand-off with named barriers: FULL[b] (loaders arrive, compute waits), DONE[b] (compute arrives when
// buffer b holds the gradient tiles, loaders wait, store them and refill the buffer).  Shared memory per pair (66.5 KB) is what
// bounds the number of pairs in flight; two buffers + one transform at a time with 16 warps keeps every
// stage short instead of interleaving three slow CTAs.
template <int P, typename T, bool LUMA3>
__global__ void __launch_bounds__(PairCfg<P>::NT, 1) pair_kernel(const __grid_constant__ Params prm) {
    using Cfg = PairCfg<P>;
    constexpr int BAR_COMPUTE = 1, BAR_FULL = 2, BAR_DONE = 4;  // ids 2,3 and 4,5; 6 = loader group
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4* buf0 = reinterpret_cast<float4*>(smem_raw);
    float4* buf1 = buf0 + P * Cfg::LD;
    float4* tw = buf1 + P * Cfg::LD;
    __shared__ float red[4 * 32];
    {
        const BlockCtx all{(int)threadIdx.x, (int)blockDim.x};
        fill_twiddles4<P>(all, tw);
        __syncthreads();
        fill_row_twiddles4<P>(all, tw, tw + P);
    }
    __syncthreads();
    const int npairs = (prm.tiles_total + 1) >> 1;
    const bool want_grad = prm.grad != nullptr;
    if (threadIdx.x >= Cfg::NT_COMPUTE) {
        // ---------------- loader / storer warps ----------------
        const GroupCtx<Cfg::NT_LOAD, 6> ctx{(int)threadIdx.x - Cfg::NT_COMPUTE, nullptr};
        int iter = 0, pr = blockIdx.x;
        for (; pr < npairs; pr += gridDim.x, ++iter) {
            const int b = iter & 1;
            float4* s = b ? buf1 : buf0;
            if (iter >= 2) {
                // buffer b holds the finished gradient tiles of pair iter-2: write them out, then reuse it
                bar_sync(BAR_DONE + b, Cfg::NT);
                if (want_grad) {
                    const int pa = 2 * (pr - 2 * (int)gridDim.x), pb = pa + 1 < prm.tiles_total ? pa + 1 : pa;
                    pair_store<P, T, LUMA3>(ctx, prm, decode_tile(prm, pa), decode_tile(prm, pb), pb != pa, s);
                    bar_sync(6, Cfg::NT_LOAD);  // all loader reads of buffer b precede its refill
                }
            }
            const int ta = 2 * pr, tb = ta + 1 < prm.tiles_total ? ta + 1 : ta;
            pair_load<P, T, LUMA3>(ctx, prm, decode_tile(prm, ta), decode_tile(prm, tb), s);
            bar_arrive(BAR_FULL + b, Cfg::NT);
        }
        // drain: the last (up to) two pairs of this CTA
        for (int back = (iter >= 2 ? 2 : iter); back >= 1; --back) {
            const int it2 = iter - back, b = it2 & 1;
            const int p2 = (int)blockIdx.x + it2 * (int)gridDim.x;
            bar_sync(BAR_DONE + b, Cfg::NT);
            if (want_grad) {
                const int pa = 2 * p2, pb = pa + 1 < prm.tiles_total ? pa + 1 : pa;
                pair_store<P, T, LUMA3>(ctx, prm, decode_tile(prm, pa), decode_tile(prm, pb), pb != pa, b ? buf1 : buf0);
            }
        }
    } else {
        // ---------------- compute warps ----------------
        GroupCtx<Cfg::NT_COMPUTE, BAR_COMPUTE> ctx{(int)threadIdx.x, nullptr};
        int iter = 0;
        for (int pr = blockIdx.x; pr < npairs; pr += gridDim.x, ++iter) {
            const int b = iter & 1;
            const int ta = 2 * pr;
            const bool b_valid = ta + 1 < prm.tiles_total;
            const int tb = b_valid ? ta + 1 : ta;
            float4* s = b ? buf1 : buf0;
            float2 accA = make_float2(0.f, 0.f), accP = make_float2(0.f, 0.f);
            ctx.trace = (prm.trace != nullptr && iter < 6) ? prm.trace + ((long long)blockIdx.x * 6 + iter) * 16 : nullptr;
            ctx.mark(0);
            bar_sync(BAR_FULL + b, Cfg::NT);  // the loaders have filled buffer b
            pair_compute<P, T, LUMA3>(ctx, prm, decode_tile(prm, ta), decode_tile(prm, tb), b_valid, s, tw, accA, accP);
            // pair_compute ends with a group barrier: buffer b now holds the gradient tiles (or is dead)
            bar_arrive(BAR_DONE + b, Cfg::NT);
            group_sum4<Cfg::NT_COMPUTE, BAR_COMPUTE>(ctx.tid, red, accA.x, accA.y, accP.x, accP.y);
            if (ctx.tid == 0) {
                prm.partials[2 * ta] = accA.x;
                prm.partials[2 * ta + 1] = accP.x;
                if (b_valid) {
                    prm.partials[2 * tb] = accA.y;
                    prm.partials[2 * tb + 1] = accP.y;
                }
                if (ctx.trace != nullptr) {
                    unsigned smid;
                    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
                    ctx.trace[14] = smid;
                    unsigned long long gt;
                    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
                    ctx.trace[15] = (long long)gt;
                }
            }
        }
    }
    finish(prm, gridDim.x);
}

// Programmatic dependent launch (sm_90+): consecutive launches of this library are chained with the
// programmatic-stream-serialization attribute, so the next grid is scheduled while the previous one drains and its
// CTAs sit in `griddepcontrol.wait` until that grid has completed and flushed -- the launch latency between
// dependent kernels (~2 us each, a few percent of a 100 us step) overlaps the tail.  Both instructions are no-ops in
// a launch without the attribute.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// Issued by every CTA when its own work is done: releasing earlier lets the dependent grid's CTAs take SM slots
// that this grid's not-yet-started CTAs need (measured: 118 us instead of 98 us per step).
__device__ __forceinline__ void pdl_release() { asm volatile("griddepcontrol.launch_dependents;"); }

// Thread-per-line kernel for 64 x 64 tiles (line_tile.cuh): 64 threads = one tile, six CTAs per SM.
template <typename T, bool LUMA3>
__global__ void __launch_bounds__(LineCfg::NT, 6) line_kernel(const __grid_constant__ Params prm) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* s = reinterpret_cast<float2*>(smem_raw);
    BlockCtxT<LineCfg::NT> ctx{(int)threadIdx.x, nullptr};
    pdl_wait();
    int iter = 0;
    for (int tile = blockIdx.x; tile < prm.tiles_total; tile += gridDim.x, ++iter) {
        float a = 0.f, p = 0.f;
        ctx.trace = (prm.trace != nullptr && iter < 6) ? prm.trace + ((long long)blockIdx.x * 6 + iter) * 16 : nullptr;
        const int nt = tile + (int)gridDim.x;
        line_process<T, LUMA3>(ctx, prm, tile, s, a, p, nt < prm.tiles_total ? nt : -1);
        block_sum2(a, p);
        if (threadIdx.x == 0) {
            prm.partials[2 * tile] = a;
            prm.partials[2 * tile + 1] = p;
            if (ctx.trace != nullptr) ctx.trace[15] = 1;
        }
    }
    pdl_release();
    finish(prm, gridDim.x);
}

// Sub-tile path, launches 1 and 3 (sub_tile.cuh): thread-per-line 64 x 64 transforms of the D x D decimated
// sub-images.  Forward: one CTA = one sub-image PAIR (adjacent pixel columns, so the source rows are read as
// 8-byte pairs), two 64-thread groups with a work tile each.  Inverse: one 64-thread CTA = one packed plane.
template <typename T, bool LUMA3>
__global__ void __launch_bounds__(SubCfg::NT_FWD, 3) sub_fwd_kernel(const __grid_constant__ Params prm) {
    pdl_wait();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* s = reinterpret_cast<float2*>(smem_raw);
    BlockCtxT<SubCfg::NT_FWD> ctx{(int)threadIdx.x, nullptr};
    const int nunits = prm.chunk_now * (prm.sub_d * prm.sub_d / 2);
    int iter = 0;
    for (int u = blockIdx.x; u < nunits; u += gridDim.x, ++iter) {
        ctx.trace = (prm.trace != nullptr && iter < 6) ? prm.trace + ((long long)blockIdx.x * 6 + iter) * 16 : nullptr;
        sub_fwd_process<T, LUMA3>(ctx, prm, u, s);
        if (ctx.trace != nullptr && threadIdx.x == 0) ctx.trace[15] = 1;
    }
    pdl_release();
}
// D = 4 forward launch as 2-CTA clusters: the two CTAs of a cluster own the two column pairs of one (tile, row phase),
// load half of the rows each with full-sector 16-byte loads and hand the other CTA its half through distributed
// shared memory (sub_fwd_load_quad).  Cluster barriers fence the hand-over in both directions.
template <typename T, bool LUMA3>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(SubCfg::NT_FWD, 3) sub_fwd4_kernel(const __grid_constant__ Params prm) {
    namespace cg = cooperative_groups;
    pdl_wait();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* s = reinterpret_cast<float2*>(smem_raw);
    cg::cluster_group cl = cg::this_cluster();
    const int rank = (int)cl.block_rank();
    float2* peer = cl.map_shared_rank(s, rank ^ 1);
    float2* dst01 = rank == 0 ? s : peer;
    float2* dst23 = rank == 0 ? peer : s;
    BlockCtxT<SubCfg::NT_FWD> ctx{(int)threadIdx.x, nullptr};
    const int npairs = prm.chunk_now * 4;  // (tile, row phase)
    cl.sync();                             // the peer's shared memory exists from here on
    for (int w = blockIdx.x >> 1; w < npairs; w += gridDim.x >> 1) {
        SubUnit su;
        su.tile_local = w >> 2;
        su.p = w & 3;
        su.i = rank;
        su.plane = su.p * 2 + rank;
        sub_fwd_load_quad<T, LUMA3>(ctx, prm, decode_tile(prm, prm.tile_base + su.tile_local), su.p, rank, dst01, dst23);
        cl.sync();
        sub_fwd_rows(ctx, s);
        ctx.sync();
        sub_fwd_cols_store(ctx, prm, su, s);
        cl.sync();  // the peer may refill my tiles only after my column pass has read them
    }
    pdl_release();
}

template <typename T, bool LUMA3>
__global__ void __launch_bounds__(SubCfg::NT_INV, 6) sub_inv_kernel(const __grid_constant__ Params prm) {
    pdl_wait();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* s = reinterpret_cast<float2*>(smem_raw);
    BlockCtxT<SubCfg::NT_INV> ctx{(int)threadIdx.x, nullptr};
    const int nunits = prm.chunk_now * (prm.sub_d * prm.sub_d / 2);
    int iter = 0;
    for (int u = blockIdx.x; u < nunits; u += gridDim.x, ++iter) {
        ctx.trace = (prm.trace != nullptr && iter < 6) ? prm.trace + ((long long)blockIdx.x * 6 + iter) * 16 : nullptr;
        sub_inv_process<T, LUMA3>(ctx, prm, u, s);
        if (ctx.trace != nullptr && threadIdx.x == 0) ctx.trace[15] = 1;
    }
    pdl_release();
}

// D = 4 inverse launch as 2-CTA clusters (the two packed planes i = 0, 1 of one (tile, row phase)): transforms as in
// sub_inv_kernel, then each CTA stores half of the rows with full 16-byte stores, reading the other column pair from
// the peer's shared memory (sub_inv_store_quad).
template <typename T, bool LUMA3>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(SubCfg::NT_INV, 6) sub_inv4_kernel(const __grid_constant__ Params prm) {
    namespace cg = cooperative_groups;
    pdl_wait();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* s = reinterpret_cast<float2*>(smem_raw);
    cg::cluster_group cl = cg::this_cluster();
    const int rank = (int)cl.block_rank();
    const float2* peer = cl.map_shared_rank(s, rank ^ 1);
    const float2* s01 = rank == 0 ? s : peer;
    const float2* s23 = rank == 0 ? peer : s;
    BlockCtxT<SubCfg::NT_INV> ctx{(int)threadIdx.x, nullptr};
    const int npairs = prm.chunk_now * 4;
    cl.sync();
    for (int w = blockIdx.x >> 1; w < npairs; w += gridDim.x >> 1) {
        SubUnit su;
        su.tile_local = w >> 2;
        su.p = w & 3;
        su.i = rank;
        su.plane = su.p * 2 + rank;
        sub_inv_cols(ctx, prm, su, s);
        ctx.sync();
        sub_inv_rows(ctx, s);
        cl.sync();  // both column pairs are ready
        sub_inv_store_quad<T, LUMA3>(ctx, prm, decode_tile(prm, prm.tile_base + su.tile_local), su.p, rank, s01, s23);
        cl.sync();  // the peer has read my tile
    }
    pdl_release();
}

// Sub-tile path, launch 2: per-position D x D butterflies, loss, spectral gradient (registers + L2 only).
template <int D>
__global__ void __launch_bounds__(kCombineThreads, 512 / kCombineThreads) combine_kernel(const __grid_constant__ Params prm) {
    pdl_wait();
    constexpr int PARTS = kCombineParts;
    const int lt = blockIdx.x / PARTS, part = blockIdx.x % PARTS;
    float a = 0.f, p = 0.f;
    float2* ws_tile = sub_plane(prm, lt, 0);
#pragma unroll 1
    for (int rep = 0; rep < kCombineRep; ++rep) {
        const int item = part * kCombineItemsPerPart + rep * kCombineThreads + (int)threadIdx.x;
        if (item < kCombineItems) combine_item<D>(prm, ws_tile, item, a, p);
    }
    block_sum2(a, p);
    if (threadIdx.x == 0) {
        const long long slot = (long long)(prm.tile_base + lt) * PARTS + part;
        prm.partials[2 * slot] = a;
        prm.partials[2 * slot + 1] = p;
    }
    pdl_release();
    finish(prm, (unsigned)prm.tiles_total * PARTS);
}

template <int P, typename T, bool LUMA3>
__global__ void __launch_bounds__(SplitCfg<P>::NT) split_rows_fwd_kernel(const __grid_constant__ Params prm) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* s = reinterpret_cast<float2*>(smem_raw);
    float2* tw = s + Split<P>::RS * (P + 1);
    const BlockCtx ctx{(int)threadIdx.x, (int)blockDim.x};
    fill_twiddles<P>(ctx, tw);  // prologue: independent of the previous grid
    ctx.sync();
    pdl_wait();
    split_rows_fwd<P, T, LUMA3>(ctx, prm, blockIdx.x / Split<P>::ROW_SLABS, blockIdx.x % Split<P>::ROW_SLABS, s, tw);
    pdl_release();
}

template <int P>
__global__ void __launch_bounds__(SplitCfg<P>::NT) split_cols_kernel(const __grid_constant__ Params prm) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* s = reinterpret_cast<float2*>(smem_raw);
    float2* tw = s + P * (2 * Split<P>::GS + 1);
    const BlockCtx ctx{(int)threadIdx.x, (int)blockDim.x};
    fill_twiddles<P>(ctx, tw);
    ctx.sync();
    pdl_wait();
    const int lt = blockIdx.x / Split<P>::PARTS, pair = blockIdx.x % Split<P>::PARTS;
    float a = 0.f, p = 0.f;
    split_cols<P>(ctx, prm, lt, pair, s, tw, a, p);
    pdl_release();
    block_sum2(a, p);
    if (threadIdx.x == 0) {
        const long long slot = (long long)(prm.tile_base + lt) * Split<P>::PARTS + pair;
        prm.partials[2 * slot] = a;
        prm.partials[2 * slot + 1] = p;
    }
    finish(prm, (unsigned)prm.tiles_total * Split<P>::PARTS);  // ticket runs across all chunks
}

template <int P, typename T, bool LUMA3>
__global__ void __launch_bounds__(SplitCfg<P>::NT) split_rows_inv_kernel(const __grid_constant__ Params prm) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* s = reinterpret_cast<float2*>(smem_raw);
    float2* tw = s + Split<P>::RS * (P + 1);
    const BlockCtx ctx{(int)threadIdx.x, (int)blockDim.x};
    fill_twiddles<P>(ctx, tw);
    ctx.sync();
    pdl_wait();
    split_rows_inv<P, T, LUMA3>(ctx, prm, blockIdx.x / Split<P>::ROW_SLABS, blockIdx.x % Split<P>::ROW_SLABS, s, tw);
    pdl_release();
}

// dst = src * host_scale * (*dev_scale); 16-byte vectors, grid-stride.
template <typename T>
__global__ void __launch_bounds__(256) grad_scale_kernel(T* __restrict__ dst, const T* __restrict__ src, long long numel,
                                                         const float* __restrict__ dev_scale, float host_scale) {
    const float sc = host_scale * (dev_scale ? __ldg(dev_scale) : 1.0f);
    constexpr int V = 16 / sizeof(T);
    const long long nvec = numel / V;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
        uint4 raw = reinterpret_cast<const uint4*>(src)[i];
        T* e = reinterpret_cast<T*>(&raw);
#pragma unroll
        for (int k = 0; k < V; ++k) e[k] = (T)((float)e[k] * sc);
        reinterpret_cast<uint4*>(dst)[i] = raw;
    }
    for (long long i = nvec * V + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < numel; i += stride)
        dst[i] = (T)((float)src[i] * sc);
}

// ---- regional 100 x 256 FFT loss (regional.cuh): one CTA per (image, channel, band), tile resident in shared memory
template <typename T, bool LUMA3>
__global__ void __launch_bounds__(RegCfg::NT, 1) regional_kernel(const __grid_constant__ Params prm) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* s = reinterpret_cast<float2*>(smem_raw);
    float2* tw = s + RegCfg::H * RegCfg::LD;
    float2* w100 = tw + RegCfg::W;
    const BlockCtx ctx{(int)threadIdx.x, (int)blockDim.x};
    fill_twiddles<RegCfg::W>(ctx, tw);
    reg_fill_w100(ctx, w100);
    ctx.sync();
    pdl_wait();
    for (int unit = blockIdx.x; unit < prm.tiles_total; unit += gridDim.x) {
        float a = 0.f, p = 0.f;
        regional_process<T, LUMA3>(ctx, prm, unit, s, tw, w100, a, p);
        block_sum2(a, p);
        if (threadIdx.x == 0) {
            prm.partials[2 * unit] = a;
            prm.partials[2 * unit + 1] = p;
        }
    }
    pdl_release();
    finish(prm, gridDim.x);
}

// ---- patch triplet loss (triplet.cuh): one lane group per patch row, persistent warps ------------------------
struct ShflReduce {
    int lpr;
    __device__ __forceinline__ float operator()(float v) const {
        for (int o = lpr >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        return v;
    }
};
template <typename T, int K>
__global__ void __launch_bounds__(kTripletThreads) triplet_kernel(const __grid_constant__ TripletParams tp) {
    const int lpr = tp.p / (4 * K), gpw = 32 / lpr;  // lanes per row, rows per warp pass (rows % gpw == 0 always)
    const int lane = threadIdx.x & 31, sub = lane / lpr, l = lane - sub * lpr;
    const long long nwarps = (long long)gridDim.x * (kTripletThreads / 32);
    const long long warp = (long long)blockIdx.x * (kTripletThreads / 32) + (threadIdx.x >> 5);
    const ShflReduce red{lpr};
    float loss = 0.f, act = 0.f;
    pdl_wait();
    constexpr int R = kTripletRowsInFlight;
    // a warp owns R * gpw CONSECUTIVE rows per pass (one contiguous window of the tensors is live at a time);
    // whole warps are in or out of range because rows % gpw == 0
    for (long long base = warp * gpw * R; base < tp.rows; base += nwarps * gpw * R) {
        TripletRow<K> tr[R];
#pragma unroll
        for (int r = 0; r < R; ++r)
            if (base + r * gpw < tp.rows) triplet_row_load<T, K>(tp, base + r * gpw + sub, l, lpr, tr[r]);
#pragma unroll
        for (int r = 0; r < R; ++r)
            if (base + r * gpw < tp.rows) triplet_row_finish<T, K>(tp, tr[r], l, lpr, red, loss, act);
    }
    pdl_release();
    block_sum2(loss, act);
    if (threadIdx.x == 0) {
        tp.partials[2 * blockIdx.x] = loss;
        tp.partials[2 * blockIdx.x + 1] = act;
    }
    // last CTA: fixed-order sum of the per-CTA partials in double, ticket left at zero
    __shared__ bool last;
    __shared__ double dred[2][32];
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        last = (atomicAdd(tp.counter, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    double a = 0.0, b = 0.0;
    for (unsigned i = threadIdx.x; i < gridDim.x; i += blockDim.x) {
        a += (double)__ldcg(tp.partials + 2 * i);
        b += (double)__ldcg(tp.partials + 2 * i + 1);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_down_sync(0xffffffffu, a, o);
        b += __shfl_down_sync(0xffffffffu, b, o);
    }
    if (lane == 0) {
        dred[0][threadIdx.x >> 5] = a;
        dred[1][threadIdx.x >> 5] = b;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        a = 0.0;
        b = 0.0;
        for (int w = 0; w < kTripletThreads / 32; ++w) {
            a += dred[0][w];
            b += dred[1][w];
        }
        triplet_outputs(tp, a, b);
        *tp.counter = 0u;
    }
}

// temperature map alone (vectorize_temps): red channel -> uint8 like ToPILImage -> table
struct TempsParams {
    const void* x;
    long long xs[4];
    int n, h;
    float* out;
    float lut[256];
};
template <typename T>
__global__ void __launch_bounds__(256) temps_kernel(const __grid_constant__ TempsParams tp) {
    const long long total4 = (long long)tp.n * tp.h * (tp.h / 4);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
        const int x4 = (int)(i % (tp.h / 4));
        const long long r = i / (tp.h / 4);
        const int y = (int)(r % tp.h), n = (int)(r / tp.h);
        float v[4];
        IO<T>::load4(static_cast<const T*>(tp.x) + n * tp.xs[0] + (long long)y * tp.xs[2] + 4 * x4, v);
        const float4 o = make_float4(tp.lut[IO<T>::quant(v[0])], tp.lut[IO<T>::quant(v[1])], tp.lut[IO<T>::quant(v[2])],
                                     tp.lut[IO<T>::quant(v[3])]);
        *reinterpret_cast<float4*>(tp.out + 4 * i) = o;
    }
}

}  // namespace tfcfft
