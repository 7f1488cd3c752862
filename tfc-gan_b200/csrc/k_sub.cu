// k_sub.cu -- sub-tile pipeline for P = 128 / 256 (sub_tile.cuh): three launches per workspace chunk
//   sub_fwd_kernel / sub_fwd4_kernel   sub-images -> sub-spectra S_pq           (HBM read of fake / real)
//   combine_kernel<D>                  S_pq -> Z -> loss, G -> packed planes   (L2 only)
//   sub_inv_kernel / sub_inv4_kernel   packed planes -> gradient pixel pairs   (HBM write of grad)
#include <cooperative_groups.h>

#include "launchers.h"
#include "sub_tile.cuh"
#include "combine8.cuh"
#include "combine_quad.cuh"
#include "sub_ring.cuh"

namespace tfcfft {

// ---- tile-granular dependencies between the three launches --------------------------------------------------------
// The launches are chained with programmatic dependent launch, but `griddepcontrol.wait` is a whole-grid barrier: at
// batch 64 the forward launch is 1.15 waves deep and 85 % of the SMs idle through its 15 us tail, then the same
// again after the combine launch.  Instead every launch releases its dependents at START (all launches are
// persistent, one wave: a waiting dependent CTA never takes a slot a primary CTA still needs), and a combine /
// inverse CTA waits only for ITS TILE: forward CTAs count finished units per tile (`fwd_done`), combine CTAs count
// finished parts (`cmb_done`), both in the workspace header with release / acquire semantics.  Combine CTAs of the
// tiles of the first wave run on the SMs the forward tail leaves idle, inverse CTAs start as tiles complete.  Only the
// FORWARD launch still waits for the whole previous grid (it overwrites the workspace the previous call's inverse
// launch reads).  NEGATIVE RESULT (profiles/r02_sub_pipe_ab.txt): bit-identical, but 10 % slower than whole-grid
// waits on every workload, so this mode is opt-in (TFCFFT_FINE_DEPS=1) and whole-grid waits are the default.
__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_add(unsigned* p, unsigned v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned* sched_fwd_done(const Params& prm) { return prm.sched + (kSchedFwdDone - kSchedHeads) / 4; }
__device__ __forceinline__ unsigned* sched_cmb_done(const Params& prm) { return prm.sched + (kSchedCmbDone - kSchedHeads) / 4; }
// all threads of the CTA: publish this CTA's global stores, then count one finished item
__device__ __forceinline__ void sched_signal(unsigned* counter) {
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) red_release_add(counter, 1u);
}
// all threads of the CTA: wait until `counter` reaches `target` (thread 0 polls)
__device__ __forceinline__ void sched_wait(const unsigned* counter, unsigned target) {
    if (threadIdx.x == 0) {
        while (ld_acquire(counter) < target) __nanosleep(100);
    }
    __syncthreads();
}
// last CTA of the LAST launch of a chunk: final reduction after the last chunk, counters back to zero
__device__ __forceinline__ void sched_finish(const Params& prm, bool reduce) {
    __shared__ bool last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        last = (atomicAdd(prm.sched + 3, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!last) return;
    if (reduce && prm.tile_base + prm.chunk_now >= prm.tiles_total) finalize_sums(prm);
    __syncthreads();
    unsigned* fd = sched_fwd_done(prm);
    unsigned* cd = sched_cmb_done(prm);
    for (int i = (int)threadIdx.x; i < prm.chunk_now; i += (int)blockDim.x) {
        fd[i] = 0u;
        cd[i] = 0u;
    }
    if (threadIdx.x == 0) prm.sched[3] = 0u;
}


// Execution-only cluster barrier for the write-after-read hand-backs ("the peer has finished READING my tile", so I may
// overwrite it / exit): cg::cluster_group::sync() arrives with release semantics, and a release in front of which the
// CTA has global stores in flight (workspace planes, gradient rows) makes every thread wait for their L2 round trip
// (ncu: MEMBAR stall, 3-6 % of the forward / inverse launches).  The peer's reads completed before it issued its
// dependent stores, so no memory ordering is needed here.  TFCFFT_NO_RELAXED_BAR builds fall back to sync().
// split form: arrive as soon as the CTA runs, wait right before the first access to a peer's shared memory ("the peer has
// started"), so the barrier's latency hides behind the whole-grid wait
#ifdef TFCFFT_NO_SPLIT_BAR  // A/B builds: one full barrier where the wait is
__device__ __forceinline__ void cluster_arrive_exec() {}
__device__ __forceinline__ void cluster_wait_exec() { cooperative_groups::this_cluster().sync(); }
#else
__device__ __forceinline__ void cluster_arrive_exec() { asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait_exec() { asm volatile("barrier.cluster.wait.aligned;" ::: "memory"); }
#endif
__device__ __forceinline__ void cluster_sync_exec() {
#ifdef TFCFFT_NO_RELAXED_BAR
    cooperative_groups::this_cluster().sync();
#else
    asm volatile("barrier.cluster.arrive.relaxed.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory");
#endif
}

// ---- L2 look-ahead of the forward launches ----------------------------------------------------------------------------
// A forward CTA alternates between a load phase (DRAM-bound) and two transform passes (DRAM idle); at batch 64 the second
// round of the 256 x 256 launch (34 of 256 row phases on 222 cluster slots) is a lone CTA streaming its rows at DRAM
// latency.  Once a CTA's own loads have landed it asks L2 for the rows of its NEXT unit (one bulk prefetch per pixel row
// and channel; plain line prefetches when a row is not 16-byte aligned), so they arrive while it transforms; the first
// unit is requested before the whole-grid wait (the inputs are not written by this library's earlier launches, and a
// prefetch cannot change results either way).  Rows [a0, a0 + na) of row phase p of the tile.
// NEGATIVE RESULT, opt-in (TFCFFT_SUB_LOOKAHEAD=1|2|3): see launch_sub.
__device__ __forceinline__ void l2_prefetch_row(const void* q, unsigned bytes) {
    const unsigned long long a = (unsigned long long)q;
    if (((a | bytes) & 15ull) == 0ull) {
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(q), "r"(bytes) : "memory");
    } else {
        for (unsigned long long l = a & ~127ull; l < a + bytes; l += 128ull) asm volatile("prefetch.global.L2 [%0];" ::"l"(l));
    }
}
template <typename T, bool LUMA3>
__device__ __forceinline__ void sub_fwd_lookahead(const Params& prm, int tile, int p, int a0, int na) {
    constexpr int NC = LUMA3 ? 3 : 1;
    const int D = prm.sub_d, P = 64 * D;
    const TileCoord tc = decode_tile(prm, tile);
    const T* fp = tile_ptr<T>(prm.fake, prm.fs, tc, P);
    const T* rp = real_tile_ptr<T>(prm, tc, P);
    const unsigned bytes = (unsigned)(P * sizeof(T));
    for (int it = (int)threadIdx.x; it < na * NC * 2; it += (int)blockDim.x) {
        const int h = it & 1, c = (it >> 1) % NC, y = D * (a0 + (it >> 1) / NC) + p;
        l2_prefetch_row(h ? rp + y * prm.rs[2] + c * prm.rs[1] : fp + y * prm.fs[2] + c * prm.fs[1], bytes);
    }
}

// ---- deferred final reduction ------------------------------------------------------------------------------------------
// finish() at the end of every combine CTA is a serial chain -- barrier, fence, ticket atomic and its round trip, barrier --
// during which the CTA holds its slot (ncu per-instruction page: 13 % of the warp time of combine_kernel<4> at one wave, 30 %
// of combine_kernel<2> at four waves).  With a gradient the combine launches are always followed by an inverse launch, so
// they only store their partial sums and the sum is taken behind a kernel boundary instead: by `fin_ctas` CTAs appended to
// the last inverse launch (one cluster's worth; the first one works, after the whole-grid wait, next to the workers), or
// -- two-lane schedule, where the last inverse launch is not ordered behind the other lane's combine launches -- by a
// one-CTA launch after the join.  TFCFFT_NO_DEFER=1 restores the ticket.
__device__ __forceinline__ bool fin_cta(const Params& prm) {  // true: this CTA is not a worker
    if ((int)blockIdx.x < (int)gridDim.x - prm.fin_ctas) return false;
    if ((int)blockIdx.x == (int)gridDim.x - prm.fin_ctas) {
        pdl_wait();
        finalize_sums(prm);
    }
    return true;
}
template <int DT>  // one instance per translation unit (k_sub.cu is compiled once per element type)
__global__ void __launch_bounds__(128) sub_finalize_kernel(const __grid_constant__ Params prm) {
    pdl_wait();
    finalize_sums(prm);
}

// Forward: one CTA = one sub-image PAIR (adjacent pixel columns, so the source rows are read as
// 8-byte pairs), two 64-thread groups with a work tile each.  Inverse: one 64-thread CTA = one packed plane.
template <typename T, bool LUMA3>
__global__ void __launch_bounds__(SubCfg::NT_FWD, 3) sub_fwd_kernel(const __grid_constant__ Params prm) {
    const int npp = prm.sub_d * prm.sub_d / 2, nunits = prm.chunk_now * npp, hd = prm.sub_d / 2;
    // look-ahead: a unit reads all 64 rows of its row phase (the column pairs of one row phase share them)
    if ((prm.lookahead & 2) && (int)blockIdx.x < nunits)
        sub_fwd_lookahead<T, LUMA3>(prm, prm.tile_base + (int)blockIdx.x / npp, ((int)blockIdx.x % npp) / hd, 0, 64);
    pdl_wait();  // whole previous grid: this launch overwrites the workspace planes
    if (prm.fine_deps) pdl_release();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* s = reinterpret_cast<float2*>(smem_raw);
    BlockCtxT<SubCfg::NT_FWD> ctx{(int)threadIdx.x, nullptr};
    int iter = 0;
    for (int u = blockIdx.x; u < nunits; u += gridDim.x, ++iter) {
        ctx.trace = (prm.trace != nullptr && iter < 6) ? prm.trace + ((long long)blockIdx.x * 6 + iter) * 16 : nullptr;
        const SubUnit su = sub_unit(u, prm.sub_d);
        ctx.mark(0);
        const TileCoord tc = decode_tile(prm, prm.tile_base + su.tile_local);
        const bool same = sub_d2_quads(prm) ? sub_fwd_load_d2<T, LUMA3>(ctx, prm, tc, su, s) : sub_fwd_load<T, LUMA3>(ctx, prm, tc, su, s);
        if ((prm.lookahead & 1) && u + (int)gridDim.x < nunits) {
            const int un = u + (int)gridDim.x;
            sub_fwd_lookahead<T, LUMA3>(prm, prm.tile_base + un / npp, (un % npp) / hd, 0, 64);
        }
        ctx.sync();
        ctx.mark(1);
#pragma unroll 1
        for (int pass = 0; pass < 2; ++pass) {  // rolled: one copy of the 64-point core
            sub_fwd_pass(ctx, prm, su, s, pass);
            ctx.sync();
            ctx.mark(2 + pass);
        }
        if (prm.eq != nullptr) {  // one "fake == real" byte per unit, rewritten by every launch
            const int all_same = __syncthreads_and(same);
            if (threadIdx.x == 0) prm.eq[u] = (unsigned char)all_same;
        }
        if (ctx.trace != nullptr && threadIdx.x == 0) ctx.trace[15] = 1;
        if (prm.fine_deps) sched_signal(sched_fwd_done(prm) + u / npp);
    }
    if (!prm.fine_deps) pdl_release();
}
// D = 4 forward launch as 2-CTA clusters: the two CTAs of a cluster own the two column pairs of one (tile, row phase),
// load half of the rows each with full-sector 16-byte loads and hand the other CTA its half through distributed
// shared memory (sub_fwd_load_quad).  Cluster barriers fence the hand-over in both directions.
template <typename T, bool LUMA3>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(SubCfg::NT_FWD, 3) sub_fwd4_kernel(const __grid_constant__ Params prm) {
    namespace cg = cooperative_groups;
    cg::cluster_group cl = cg::this_cluster();
    const int rank = (int)cl.block_rank();
    cluster_arrive_exec();
    // look-ahead: this CTA reads rows [32 rank, 32 rank + 32) of row phase w & 3 of tile w >> 2
    if ((prm.lookahead & 2) && (int)(blockIdx.x >> 1) < prm.chunk_now * 4)
        sub_fwd_lookahead<T, LUMA3>(prm, prm.tile_base + (int)(blockIdx.x >> 3), (int)(blockIdx.x >> 1) & 3, 32 * rank, 32);
    pdl_wait();  // whole previous grid: this launch overwrites the workspace planes
    if (prm.fine_deps) pdl_release();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* s = reinterpret_cast<float2*>(smem_raw);
    float2* peer = cl.map_shared_rank(s, rank ^ 1);
    float2* dst01 = rank == 0 ? s : peer;
    float2* dst23 = rank == 0 ? peer : s;
    BlockCtxT<SubCfg::NT_FWD> ctx{(int)threadIdx.x, nullptr};
    const int npairs = prm.chunk_now * 4;  // (tile, row phase)
    cluster_wait_exec();                   // the peer's shared memory exists from here on
    for (int w = blockIdx.x >> 1; w < npairs; w += gridDim.x >> 1) {
        SubUnit su;
        su.tile_local = w >> 2;
        su.p = w & 3;
        su.i = rank;
        su.plane = su.p * 2 + rank;
        const bool same = sub_fwd_load_quad<T, LUMA3>(ctx, prm, decode_tile(prm, prm.tile_base + su.tile_local), su.p, rank, dst01, dst23);
        if ((prm.lookahead & 1) && w + (int)(gridDim.x >> 1) < npairs) {
            const int wn = w + (int)(gridDim.x >> 1);
            sub_fwd_lookahead<T, LUMA3>(prm, prm.tile_base + (wn >> 2), wn & 3, 32 * rank, 32);
        }
        // "fake == real" on the half of the row phase this CTA loaded.  The flag byte is stored AFTER the last cluster
        // barrier of the item: a global store in front of a cluster barrier (release semantics) makes every thread
        // wait for its L2 round trip (measured: 2 us per step)
        const int all_same = prm.eq != nullptr ? __syncthreads_and(same) : 0;
        cl.sync();
#pragma unroll 1
        for (int pass = 0; pass < 2; ++pass) {  // rolled: one copy of the 64-point core
            sub_fwd_pass(ctx, prm, su, s, pass);
            if (pass == 0) ctx.sync();
        }
        if (prm.fine_deps) sched_signal(sched_fwd_done(prm) + su.tile_local);  // this CTA's two planes are out
        cluster_sync_exec();  // the peer may refill my tiles only after my column pass has read them
        if (prm.eq != nullptr && threadIdx.x == 0) prm.eq[w * 2 + rank] = (unsigned char)all_same;
    }
    if (!prm.fine_deps) pdl_release();
}

// D = 8 forward launch as 4-CTA clusters: the four CTAs of a cluster own the four column pairs of one (tile, row
// phase), load a quarter of the rows each with full-sector loads and hand the other CTAs their columns through
// distributed shared memory (sub_fwd_load_oct).
template <typename T, bool LUMA3>
__global__ void __cluster_dims__(4, 1, 1) __launch_bounds__(SubCfg::NT_FWD, 3) sub_fwd8_kernel(const __grid_constant__ Params prm) {
    namespace cg = cooperative_groups;
    cg::cluster_group cl = cg::this_cluster();
    const int rank = (int)cl.block_rank();
    cluster_arrive_exec();
    // look-ahead: this CTA reads rows [16 rank, 16 rank + 16) of row phase w & 7 of tile w >> 3
    if ((prm.lookahead & 2) && (int)(blockIdx.x >> 2) < prm.chunk_now * 8)
        sub_fwd_lookahead<T, LUMA3>(prm, prm.tile_base + (int)(blockIdx.x >> 5), (int)(blockIdx.x >> 2) & 7, 16 * rank, 16);
    pdl_wait();  // whole previous grid: this launch overwrites the workspace planes
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* s = reinterpret_cast<float2*>(smem_raw);
    float2* const dst[4] = {cl.map_shared_rank(s, 0), cl.map_shared_rank(s, 1), cl.map_shared_rank(s, 2), cl.map_shared_rank(s, 3)};
    BlockCtxT<SubCfg::NT_FWD> ctx{(int)threadIdx.x, nullptr};
    const int nrows = prm.chunk_now * 8;  // (tile, row phase)
    cluster_wait_exec();                  // the peers' shared memory exists from here on
    for (int w = blockIdx.x >> 2; w < nrows; w += gridDim.x >> 2) {
        SubUnit su;
        su.tile_local = w >> 3;
        su.p = w & 7;
        su.i = rank;
        su.plane = su.p * 4 + rank;
        const bool same = sub_fwd_load_oct<T, LUMA3>(ctx, prm, decode_tile(prm, prm.tile_base + su.tile_local), su.p, rank, dst);
        if ((prm.lookahead & 1) && w + (int)(gridDim.x >> 2) < nrows) {
            const int wn = w + (int)(gridDim.x >> 2);
            sub_fwd_lookahead<T, LUMA3>(prm, prm.tile_base + (wn >> 3), wn & 7, 16 * rank, 16);
        }
        // "fake == real" on the quarter of the row phase this CTA loaded (stored after the item's last cluster barrier)
        const int all_same = prm.eq != nullptr ? __syncthreads_and(same) : 0;
        cl.sync();
#pragma unroll 1
        for (int pass = 0; pass < 2; ++pass) {  // rolled: one copy of the 64-point core
            sub_fwd_pass(ctx, prm, su, s, pass);
            if (pass == 0) ctx.sync();
        }
        cluster_sync_exec();  // the peers may refill my tiles only after my column pass has read them
        if (prm.eq != nullptr && threadIdx.x == 0) prm.eq[w * 4 + rank] = (unsigned char)all_same;
    }
    pdl_release();
}

template <typename T, bool LUMA3>
__global__ void __launch_bounds__(SubCfg::NT_INV, 6) sub_inv_kernel(const __grid_constant__ Params prm) {
    if (prm.fin_ctas && fin_cta(prm)) return;
    const int nworkers = (int)gridDim.x - prm.fin_ctas;
    if (prm.fine_deps) pdl_release();
    else pdl_wait();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* s = reinterpret_cast<float2*>(smem_raw);
    BlockCtxT<SubCfg::NT_INV> ctx{(int)threadIdx.x, nullptr};
    const int npp = prm.sub_d * prm.sub_d / 2, nunits = prm.chunk_now * npp;
    int iter = 0;
    for (int u = blockIdx.x; u < nunits; u += nworkers, ++iter) {
        ctx.trace = (prm.trace != nullptr && iter < 6) ? prm.trace + ((long long)blockIdx.x * 6 + iter) * 16 : nullptr;
        if (prm.fine_deps) sched_wait(sched_cmb_done(prm) + u / npp, (unsigned)kCombineParts);
        sub_inv_process<T, LUMA3>(ctx, prm, u, s);
        if (ctx.trace != nullptr && threadIdx.x == 0) ctx.trace[15] = 1;
    }
    if (prm.fine_deps) sched_finish(prm, true);
    else pdl_release();
}

// D = 4 inverse launch as 2-CTA clusters (the two packed planes i = 0, 1 of one (tile, row phase)): transforms as in
// sub_inv_kernel, then each CTA stores half of the rows with full 16-byte stores, reading the other column pair from
// the peer's shared memory (sub_inv_store_quad).
template <typename T, bool LUMA3>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(SubCfg::NT_INV, 6) sub_inv4_kernel(const __grid_constant__ Params prm) {
    namespace cg = cooperative_groups;
    if (prm.fin_ctas && fin_cta(prm)) return;  // a whole cluster: none of its CTAs reaches a cluster barrier
    const int nworkers = (int)gridDim.x - prm.fin_ctas;
    if (prm.fine_deps) pdl_release();
    else pdl_wait();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* s = reinterpret_cast<float2*>(smem_raw);
    cg::cluster_group cl = cg::this_cluster();
    const int rank = (int)cl.block_rank();
    const float2* peer = cl.map_shared_rank(s, rank ^ 1);
    BlockCtxT<SubCfg::NT_INV> ctx{(int)threadIdx.x, nullptr};
    const int npairs = prm.chunk_now * 4;
    // no barrier here: the first access to the peer's tile is behind the cluster barrier of the first item
    for (int w = blockIdx.x >> 1; w < npairs; w += nworkers >> 1) {
        SubUnit su;
        su.tile_local = w >> 2;
        su.p = w & 3;
        su.i = rank;
        su.plane = su.p * 2 + rank;
        if (prm.fine_deps) sched_wait(sched_cmb_done(prm) + su.tile_local, (unsigned)kCombineParts);
#pragma unroll 1
        for (int pass = 0; pass < 2; ++pass) {  // rolled: one copy of the 64-point core
            sub_inv_pass(ctx, prm, su, s, pass);
            if (pass == 0) ctx.sync();
        }
        cl.sync();  // both column pairs are ready
        sub_inv_store_quad<T, LUMA3>(ctx, prm, decode_tile(prm, prm.tile_base + su.tile_local), su.p, rank, s, peer);
        cluster_sync_exec();  // the peer has read my tile
    }
    if (prm.fine_deps) sched_finish(prm, true);
    else pdl_release();
}

// D = 4 inverse launch WITHOUT clusters (the product path): one 128-thread CTA = one (tile, row phase) = both packed planes,
// one 64-thread group per plane (group-local barrier between the two passes), then all 128 threads store the 64
// gradient rows of the row phase with full 16-byte stores from local shared memory (sub_inv_store_rows4).  The 2-CTA
// cluster form above pulled the other column pair out of the peer's shared memory: ncu showed a third of that launch's
// warp time waiting on those reads (8 exposed round trips per CTA at ~5 B / clk / SM of distributed-shared-memory
// bandwidth).  Same threads per SM (256 CTAs x 128 instead of 512 x 64), same arithmetic, bit-identical gradient.
template <typename T, bool LUMA3>
__global__ void __launch_bounds__(SubCfg::NT_FWD, 3) sub_inv4w_kernel(const __grid_constant__ Params prm) {
    if (prm.fin_ctas && fin_cta(prm)) return;
    const int nworkers = (int)gridDim.x - prm.fin_ctas;
    pdl_wait();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* s = reinterpret_cast<float2*>(smem_raw);
    const int g = (int)threadIdx.x >> 6;
    float2* sg = s + g * 64 * SubCfg::LD;
    BlockCtxT<64> cg{(int)threadIdx.x & 63, nullptr};
    BlockCtxT<SubCfg::NT_FWD> cb{(int)threadIdx.x, nullptr};
    const int npairs = prm.chunk_now * 4;
    for (int w = blockIdx.x; w < npairs; w += nworkers) {
        SubUnit su;
        su.tile_local = w >> 2;
        su.p = w & 3;
        su.i = g;
        su.plane = su.p * 2 + g;
#pragma unroll 1
        for (int pass = 0; pass < 2; ++pass) {  // rolled: one copy of the 64-point core
            sub_inv_pass(cg, prm, su, sg, pass);
            if (pass == 0) bar_sync(1 + g, 64);
        }
        __syncthreads();  // both planes are ready
        sub_inv_store_rows4<T, LUMA3>(cb, prm, decode_tile(prm, prm.tile_base + su.tile_local), su.p, s, s + 64 * SubCfg::LD);
        __syncthreads();  // before the next item's column pass overwrites the tiles
    }
    pdl_release();
}

// D = 8 inverse launch as 4-CTA clusters (the four packed planes i = 0..3 of one (tile, row phase)): transforms as
// in sub_inv_kernel, then each CTA stores a quarter of the rows with full-sector stores, reading the other column
// pairs from the peers' shared memory (sub_inv_store_oct).
template <typename T, bool LUMA3>
__global__ void __cluster_dims__(4, 1, 1) __launch_bounds__(SubCfg::NT_INV, 6) sub_inv8_kernel(const __grid_constant__ Params prm) {
    namespace cg = cooperative_groups;
    if (prm.fin_ctas && fin_cta(prm)) return;  // a whole cluster: none of its CTAs reaches a cluster barrier
    const int nworkers = (int)gridDim.x - prm.fin_ctas;
    pdl_wait();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* s = reinterpret_cast<float2*>(smem_raw);
    cg::cluster_group cl = cg::this_cluster();
    const int rank = (int)cl.block_rank();
    const float2* const src[4] = {cl.map_shared_rank(s, 0), cl.map_shared_rank(s, 1), cl.map_shared_rank(s, 2), cl.map_shared_rank(s, 3)};
    BlockCtxT<SubCfg::NT_INV> ctx{(int)threadIdx.x, nullptr};
    const int nrows = prm.chunk_now * 8;
    // no barrier here: the first access to the peers' tiles is behind the cluster barrier of the first item
    for (int w = blockIdx.x >> 2; w < nrows; w += nworkers >> 2) {
        SubUnit su;
        su.tile_local = w >> 3;
        su.p = w & 7;
        su.i = rank;
        su.plane = su.p * 4 + rank;
#pragma unroll 1
        for (int pass = 0; pass < 2; ++pass) {  // rolled: one copy of the 64-point core
            sub_inv_pass(ctx, prm, su, s, pass);
            if (pass == 0) ctx.sync();
        }
        cl.sync();  // all four column pairs are ready
        sub_inv_store_oct<T, LUMA3>(ctx, prm, decode_tile(prm, prm.tile_base + su.tile_local), su.p, rank, src);
        cluster_sync_exec();  // the peers have read my tile
    }
    pdl_release();
}

#if TFC_DT == 0
// Launch 2: per-position D x D butterflies, loss, spectral gradient (registers + L2 only).
template <int D>
__global__ void __launch_bounds__(kCombineThreads, 512 / kCombineThreads) combine_kernel(const __grid_constant__ Params prm) {
    constexpr int PARTS = kCombineParts;
    const int lt = blockIdx.x / PARTS, part = blockIdx.x % PARTS;
    if (prm.fine_deps) {
        pdl_release();
        sched_wait(sched_fwd_done(prm) + lt, (unsigned)(D * D / 2));
    } else {
        pdl_wait();
    }
    float a = 0.f, p = 0.f;
    float2* ws_tile = sub_plane(prm, lt, 0);
    // "fake == real on this tile" flags of the forward launch
    const unsigned char* eqf = prm.eq != nullptr ? prm.eq + (long long)lt * (D * D / 2) : nullptr;
#pragma unroll 1
    for (int rep = 0; rep < kCombineRep; ++rep) {
        const int item = part * kCombineItemsPerPart + rep * kCombineThreads + (int)threadIdx.x;
        if (item < kCombineItems) combine_item<D>(prm, ws_tile, item, a, p, eqf);
    }
    block_sum2(a, p);
    if (threadIdx.x == 0) {
        const long long slot = (long long)(prm.tile_base + lt) * PARTS + part;
        prm.partials[2 * slot] = a;
        prm.partials[2 * slot + 1] = p;
    }
    if (prm.fine_deps) {
        sched_signal(sched_cmb_done(prm) + lt);
        // with a gradient the inverse launch is the last one of the chunk: it reduces and resets (sched_finish)
        if (prm.grad == nullptr) sched_finish(prm, true);
    } else {
        pdl_release();
        if (!prm.defer_finish) finish(prm, (unsigned)prm.tiles_total * PARTS);
    }
}
// Launch 2 for 256 x 256 tiles, one position pair per thread quad (combine_quad.cuh).  The first `chunk_now` CTAs run the 66
// pairs of the self-conjugate columns of one tile each on the one-thread item (first in the grid: their threads have the
// longest instruction stream), the others kCombineQRows rows of the position grid of one tile each.
#ifndef TFC_CQ_MINB
#define TFC_CQ_MINB 5
#endif
__global__ void __launch_bounds__(CombineQCfg::NT, TFC_CQ_MINB) combine_quad_kernel(const __grid_constant__ Params prm) {
    constexpr int PARTS = CombineQCfg::PARTS;
    __shared__ float4 xbuf[(CombineQCfg::NT / 32) * CombineQCfg::XWARP];
    const int bx = (int)blockIdx.x, nt = prm.chunk_now;
    const int lt = bx < nt ? bx : (bx - nt) / CombineQCfg::GROUPS, part = bx < nt ? 0 : 1 + (bx - nt) % CombineQCfg::GROUPS;
    pdl_wait();
    float a = 0.f, p = 0.f;
    float2* ws_tile = sub_plane(prm, lt, 0);
    const unsigned char* eqf = prm.eq != nullptr ? prm.eq + (long long)lt * 8 : nullptr;
    if (part == 0) {
        const int item = 64 * 32 + (int)threadIdx.x;
        if (item < kCombineItems) combine_item<4>(prm, ws_tile, item, a, p, eqf);
    } else {
        const int warp = (int)threadIdx.x >> 5, lane = (int)threadIdx.x & 31;
        const int kx0 = warp * 8 + (lane >> 2), kx = kx0 ? kx0 : 1;
        const int row0 = (part - 1) * CombineQCfg::RPC;
        unsigned long long flags = 0ull;
        if (eqf != nullptr) flags = __ldcg(reinterpret_cast<const unsigned long long*>(eqf));
        float2 sa[4], sb[4];
        combine_quad_load(ws_tile, row0, kx, sa, sb);
        const c2 wx = quad_twiddle(kx);
#pragma unroll 1
        for (int i = 0; i < CombineQCfg::RPC; ++i) {
            float2 na[4], nb[4];  // the next row's sub-spectra travel while this row is processed (last row: reloads itself)
            combine_quad_load(ws_tile, row0 + (i + 1 < CombineQCfg::RPC ? i + 1 : i), kx, na, nb);
            combine_quad_item(prm, ws_tile, row0 + i, kx, kx0 != 0, wx, flags, sa, sb, xbuf + warp * CombineQCfg::XWARP, a, p);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                sa[q] = na[q];
                sb[q] = nb[q];
            }
        }
    }
    pdl_release();
    block_sum2(a, p);
    if (threadIdx.x == 0) {
        const long long slot = (long long)(prm.tile_base + lt) * PARTS + part;
        prm.partials[2 * slot] = a;
        prm.partials[2 * slot + 1] = p;
    }
    if (!prm.defer_finish) finish(prm, (unsigned)prm.tiles_total * PARTS);
}
#ifndef TFC_C8_MINB
#define TFC_C8_MINB 2
#endif
// Launch 2 for 512 x 512 tiles: one CTA per (tile, row of the position grid), staged in shared memory (combine8.cuh).
__global__ void __launch_bounds__(Combine8Cfg::NT, TFC_C8_MINB) combine8_kernel(const __grid_constant__ Params prm) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* sm = reinterpret_cast<float2*>(smem_raw);
    constexpr int PARTS = kCombine8Parts;
    const int lt = blockIdx.x / PARTS, row = blockIdx.x % PARTS;
    pdl_wait();
    float a = 0.f, p = 0.f;
    const BlockCtx ctx{(int)threadIdx.x, Combine8Cfg::NT};
    combine8_rows(ctx, prm, sub_plane(prm, lt, 0), row, sm, a, p, prm.eq != nullptr ? prm.eq + (long long)lt * 32 : nullptr);
    pdl_release();
    block_sum2(a, p);
    if (threadIdx.x == 0) {
        const long long slot = (long long)(prm.tile_base + lt) * PARTS + row;
        prm.partials[2 * slot] = a;
        prm.partials[2 * slot + 1] = p;
    }
    if (!prm.defer_finish) finish(prm, (unsigned)prm.tiles_total * PARTS);
}
// auxiliary stream + fork / join events of the two-lane schedule (launch_sub), one set per device, created on first use
Lanes* lanes_get() {
    static Lanes table[kMaxDevices];
    static std::atomic<int> ready[kMaxDevices];
    static std::mutex mu;
    const int dev = current_device();
    if (!ready[dev].load(std::memory_order_acquire)) {
        std::lock_guard<std::mutex> lk(mu);
        if (!ready[dev].load(std::memory_order_relaxed)) {
            Lanes& l = table[dev];
            if (cudaStreamCreateWithFlags(&l.aux, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
            if (cudaEventCreateWithFlags(&l.fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
            if (cudaEventCreateWithFlags(&l.join, cudaEventDisableTiming) != cudaSuccess) return nullptr;
            ready[dev].store(1, std::memory_order_release);
        }
    }
    return &table[dev];
}
cudaError_t launch_combine(int d, int grid, const Params& prm, cudaStream_t st) {
    if (d == 8) return launch_pdl(combine8_kernel, grid, Combine8Cfg::NT, Combine8Cfg::SMEM, st, prm);
    if (d == 4 && combine_quad_enabled(256, prm.flags)) return launch_pdl(combine_quad_kernel, grid, CombineQCfg::NT, 0, st, prm);
    return d == 2 ? launch_pdl(combine_kernel<2>, grid, kCombineThreads, 0, st, prm)
                  : launch_pdl(combine_kernel<4>, grid, kCombineThreads, 0, st, prm);
}
#endif

// ---------------------------------------------------------------------------------------------------------------
// Pipelined variant: ONE persistent launch per workspace chunk instead of three.  The three launches above are each
// about one wave deep at batch 64 (1.15 / 1 / 1.15 waves): three wave tails, no overlap between the HBM-bound loads of
// launch 1, the L2 / issue-bound combine and the HBM-bound stores of launch 3.  Here every CTA pulls work items from
// three statically partitioned queues -- inverse (gradient store), combine, forward (pixel load) -- and a tile's combine items
// become ready when its D*D/2 forward items have signalled, its inverse items when its 9 combine parts have: tile
// j's loads overlap tile i's combine and tile h's stores on the same SM, and the only tail is the last tile's.
//   * items run the SAME device functions as the three launches (sub_fwd_process, combine_item, sub_inv_process):
//     results are bit-identical to the 3-launch path (tests compare them);
//   * no deadlock: forward items wait for nothing, a combine / inverse item is only CLAIMED when its dependencies
//     have completed, and all CTAs are co-resident (grid = SMs x occupancy);
//   * scheduler state (per-tile done counters, exit ticket) lives in the zero-initialised workspace header and is
//     reset by the last CTA to leave, like the finalise ticket.
struct PipeCfg {
    static constexpr int NT = 128;
    static constexpr size_t SMEM = SubCfg::SMEM_FWD;  // two work tiles: a forward pair, or two inverse planes
};

template <typename T, bool LUMA3>
__global__ void __launch_bounds__(PipeCfg::NT, 3) sub_pipe_kernel(const __grid_constant__ Params prm) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* s = reinterpret_cast<float2*>(smem_raw);
    __shared__ int s_type, s_idx;
    const int tid = (int)threadIdx.x;
    const int D = prm.sub_d, npp = D * D / 2, T_ = prm.chunk_now;
    const int n_fwd = T_ * npp, n_cmb = T_ * kCombineParts;
    const int inv_per_tile = npp / 2;  // a CTA runs two inverse planes at once (two 64-thread groups)
    const int n_inv = prm.grad != nullptr ? T_ * inv_per_tile : 0;
    unsigned* heads = prm.sched;                  // [0] fwd, [1] cmb, [2] inv, [3] exit ticket
    unsigned* fwd_done = prm.sched + (kSchedFwdDone - kSchedHeads) / 4;
    unsigned* cmb_done = prm.sched + (kSchedCmbDone - kSchedHeads) / 4;
    pdl_wait();
    // Items are OWNED statically (item i of a queue belongs to CTA i mod gridDim): no claim traffic -- a first version
    // with three shared atomic queue heads spent ~0.45 us per claim serialising 444 CTAs on three L2 addresses (5 x
    // slower than the three launches).  A CTA only decides WHICH of its own next items to run: inverse if that tile's
    // combine parts are done, else combine if that tile's forward items are done, else its next forward item (never
    // waits), else it polls (only per-tile counters, read by a handful of CTAs each).
    const int G_ = (int)gridDim.x, b_ = (int)blockIdx.x;
    int nf = b_, nc = b_, ni = b_;  // this CTA's next forward / combine / inverse item
    for (;;) {
        if (tid == 0) {
            int type = 0, idx = 0;  // 0 retry, 1 fwd, 2 cmb, 3 inv, 4 exit
            if (ni < n_inv && ld_acquire(cmb_done + ni / inv_per_tile) == (unsigned)kCombineParts) {
                type = 3;
                idx = ni;
                ni += G_;
            } else if (nc < n_cmb && ld_acquire(fwd_done + nc / kCombineParts) == (unsigned)npp) {
                type = 2;
                idx = nc;
                nc += G_;
            } else if (nf < n_fwd) {
                type = 1;
                idx = nf;
                nf += G_;
            } else if (ni >= n_inv && nc >= n_cmb) {
                type = 4;
            } else {
                __nanosleep(100);
            }
            s_type = type;
            s_idx = idx;
        }
        __syncthreads();
        const int type = s_type, idx = s_idx;
        __syncthreads();  // s_type / s_idx may be rewritten by thread 0 from here on
        if (type == 4) break;
        if (type == 0) continue;
        if (type == 1) {
            const BlockCtxT<PipeCfg::NT> ctx{tid, nullptr};
            (void)sub_fwd_process<T, LUMA3>(ctx, prm, idx, s);  // ends with a block barrier
            __threadfence();
            __syncthreads();
            if (tid == 0) red_release_add(fwd_done + idx / npp, 1u);
        } else if (type == 2) {
            const int lt = idx / kCombineParts, part = idx % kCombineParts;
            float a = 0.f, p = 0.f;
            float2* ws_tile = sub_plane(prm, lt, 0);
#pragma unroll 1
            for (int rep = 0; rep < kCombineRep; ++rep) {
                const int item = part * kCombineItemsPerPart + rep * kCombineThreads + tid;
                if (item < kCombineItems) {
                    if (D == 2) combine_item<2>(prm, ws_tile, item, a, p);
                    else combine_item<4>(prm, ws_tile, item, a, p);
                }
            }
            block_sum2(a, p);
            if (tid == 0) {
                const long long slot = (long long)(prm.tile_base + lt) * kCombineParts + part;
                prm.partials[2 * slot] = a;
                prm.partials[2 * slot + 1] = p;
            }
            __threadfence();
            __syncthreads();
            if (tid == 0) red_release_add(cmb_done + lt, 1u);
        } else {
            // two packed planes of one tile, one per 64-thread group (named barriers 1 and 2)
            const int g = tid >> 6;
            const GroupCtx<64, 1> c0{tid & 63, nullptr};
            const GroupCtx<64, 2> c1{tid & 63, nullptr};
            const int u = (idx / inv_per_tile) * npp + (idx % inv_per_tile) * 2 + g;
            if (g == 0) sub_inv_process<T, LUMA3>(c0, prm, u, s);
            else sub_inv_process<T, LUMA3>(c1, prm, u, s + 64 * SubCfg::LD);
            __syncthreads();
        }
    }
    pdl_release();
    // last CTA to leave: final reduction (after the last chunk) and scheduler reset for the next launch
    __shared__ bool last;
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        last = (atomicAdd(heads + 3, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!last) return;
    if (prm.tile_base + T_ >= prm.tiles_total) finalize_sums(prm);
    __syncthreads();
    for (int i = tid; i < T_; i += PipeCfg::NT) {
        fwd_done[i] = 0u;
        cmb_done[i] = 0u;
    }
    if (tid < 4) heads[tid] = 0u;
}

namespace {

template <typename T, bool LUMA3>
int launch_sub(Params prm, cudaStream_t st) {
    auto kf = sub_fwd_kernel<T, LUMA3>;
    auto ki = sub_inv_kernel<T, LUMA3>;
    auto kf4 = sub_fwd4_kernel<T, LUMA3>;
    auto ki4 = sub_inv4_kernel<T, LUMA3>;
    auto kf8 = sub_fwd8_kernel<T, LUMA3>;
    auto ki8 = sub_inv8_kernel<T, LUMA3>;
    auto ki4w = sub_inv4w_kernel<T, LUMA3>;
    static KernelFacts ff, fi, ff4, fi4, ff8, fi8, fi4w;
    int per_sm_f = 1, per_sm_i = 1;
    if (int rc = ff.get(kf, SubCfg::NT_FWD, SubCfg::SMEM_FWD, &per_sm_f)) return rc;
    if (int rc = fi.get(ki, SubCfg::NT_INV, SubCfg::SMEM_INV, &per_sm_i)) return rc;
    const int D = prm.sub_d, npp = D * D / 2;
    static const bool no_cluster = getenv("TFCFFT_NO_CLUSTER") != nullptr;
    const bool cluster = D == 4 && !no_cluster;
    const bool cluster8 = D == 8 && !no_cluster;
    if (cluster8) {
        if (int rc = ff8.get(kf8, SubCfg::NT_FWD, SubCfg::SMEM_FWD, nullptr)) return rc;
        if (int rc = fi8.get(ki8, SubCfg::NT_INV, SubCfg::SMEM_INV, nullptr)) return rc;
    }
    int per_sm_i4w = 1;
    if (cluster) {
        if (int rc = ff4.get(kf4, SubCfg::NT_FWD, SubCfg::SMEM_FWD, nullptr)) return rc;
        if (int rc = fi4.get(ki4, SubCfg::NT_INV, SubCfg::SMEM_INV, nullptr)) return rc;
        if (int rc = fi4w.get(ki4w, SubCfg::NT_FWD, SubCfg::SMEM_FWD, &per_sm_i4w)) return rc;
    }
    // measured (profiles/r02_sub_pipe_ab.txt): the pipelined kernel is bit-identical but SLOWER than the three launches
    // (global 256^2 b64: 118 vs 73 us; 4-patch b256: 372 vs 248 us) -- one kernel that holds the forward, combine and
    // inverse code of both decimations thrashes the instruction cache, runs the combine at 12 instead of 16 warps per
    // SM and loses the 2-CTA cluster loads -- so it is opt-in (TFCFFT_SUB_PIPE=1) and the three launches stay.
    static const bool pipe = getenv("TFCFFT_SUB_PIPE") != nullptr;
    static const bool no_eq = getenv("TFCFFT_NO_EQ") != nullptr;  // A/B: identical-tile tracking off
    unsigned char* const eq0 = (!no_eq && (long long)prm.chunk_tiles * npp <= kEqFlagBytes) ? prm.eq : nullptr;
    prm.eq = nullptr;  // only the three-launch product path below tracks identical tiles
    if (pipe && D != 8) {
        auto kp = sub_pipe_kernel<T, LUMA3>;
        static KernelFacts fp;
        int per_sm = 1;
        if (int rc = fp.get(kp, PipeCfg::NT, PipeCfg::SMEM, &per_sm)) return rc;
        const int cap = device_sms() * per_sm;  // every CTA must be resident: items wait on each other
        for (int base = 0; base < prm.tiles_total; base += prm.chunk_tiles) {
            prm.tile_base = base;
            prm.chunk_now = prm.tiles_total - base < prm.chunk_tiles ? prm.tiles_total - base : prm.chunk_tiles;
            const int items = prm.chunk_now * npp;
            const int grid = items < cap ? items : cap;
            if (cudaError_t e = launch_pdl(kp, grid, PipeCfg::NT, PipeCfg::SMEM, st, prm)) return (int)e;
            g_launches++;
        }
        return 0;
    }
    // measured: tile-granular dependencies (below) are bit-identical but 10 % SLOWER than whole-grid waits on every
    // sub-tile workload (global 256^2 b64: 83 vs 75 us) -- dependent CTAs that become resident early spin on their tile's
    // counter next to the CTAs that do the work -- so they are opt-in (TFCFFT_FINE_DEPS=1)
    static const bool fine = getenv("TFCFFT_FINE_DEPS") != nullptr && getenv("TFCFFT_NO_PDL") == nullptr;
    prm.fine_deps = (fine && D != 8) ? 1 : 0;
    // measured (profiles/r02_final_ab.txt): next-unit look-ahead +-0.5 % on every sub-tile workload (the second round of a
    // launch is bound by the lone CTA's transforms, not by its loads), look-ahead before the wait -4 ... -8 %: off
    static const int lookahead = getenv("TFCFFT_SUB_LOOKAHEAD") ? atoi(getenv("TFCFFT_SUB_LOOKAHEAD")) : 0;
    prm.lookahead = lookahead;
    static const bool no_defer = getenv("TFCFFT_NO_DEFER") != nullptr;
    // measured (profiles/r02_final_ab.txt): +5.6 % global 256^2 b64, +4 % per-channel RGB, +6.8 % 4-patch b256, +4.8 % patch-16
    // at 512^2; -0.6 % at D = 8 (few large combine CTAs, and the extra launch after the join): not for D = 8
    prm.defer_finish = (prm.grad != nullptr && !prm.fine_deps && !no_defer && D != 8) ? 1 : 0;
    prm.fin_ctas = 0;
    const int sms = device_sms();
    // trim the workspace chunk to a whole number of waves of the forward launch on THIS device (smallest tile count
    // whose units fill whole waves: 222 tiles of 128 x 128 / 111 tiles of 256 x 256 on a 148-SM part at 3 CTAs per SM)
    int chunk = prm.chunk_tiles;
    {
        const int wave_units = sms * per_sm_f;
        int a = wave_units, b = npp;
        while (b) { const int t = a % b; a = b; b = t; }
        const int wave_tiles = wave_units / a;
        if (chunk >= wave_tiles) chunk = (chunk / wave_tiles) * wave_tiles;
    }
    // Two lanes: when the batch is more than two waves of the forward launch, chunks of two waves alternate between
    // the caller's stream and an auxiliary stream of this library (fork / join with events), each lane owning one half
    // of the workspace: the ragged last wave and the load-only / store-only phases of one lane's launches run under
    // the other lane's launches instead of in front of them.
    Lanes* lanes = nullptr;
    int lane_tiles = 0;
    {
        // measured (profiles/r02_d8_ab.txt): +16 % at D = 8 (512^2 global b32), +4.6 % at D = 2 (4-patch b256,
        // patch-16 at 512^2), -4 % at D = 4 (256^2 global rgb, and -9 % at b64 luma where the second lane is a short
        // chain of three latency-bound launches): two lanes unless D = 4, chunks of two waves
        static const int mode_env = getenv("TFCFFT_SUB_LANES") ? atoi(getenv("TFCFFT_SUB_LANES")) : 0;
        static const int waves = getenv("TFCFFT_SUB_WAVES") ? atoi(getenv("TFCFFT_SUB_WAVES")) : 2;
        static const int lane_tiles_env = getenv("TFCFFT_SUB_LANE_TILES") ? atoi(getenv("TFCFFT_SUB_LANE_TILES")) : 0;  // A/B: even splits
        const int mode = mode_env ? mode_env : (D == 4 ? 1 : 2);
        const int wave_tiles = lane_tiles_env > 0 ? lane_tiles_env
                                                  : (sms * per_sm_f / npp) * (waves < 1 ? 1 : waves);  // whole tiles that fit the wave(s)
        // every tile has its own workspace slot when the batch fits one workspace chunk; otherwise the lanes take
        // one half of the workspace each
        const int half = prm.tiles_total <= prm.chunk_tiles ? prm.tiles_total : prm.chunk_tiles / 2;
        if (mode >= 2 && !prm.fine_deps && wave_tiles >= 1 && prm.tiles_total > wave_tiles && half >= 1) {
            // (trimming the lane's share to whole waves was measured too: 4-patch b256 in 222-tile chunks 1.02 M vs
            // 1.04 M images/s in four even 256-tile chunks -- an extra ragged chunk costs more than ragged waves)
            lane_tiles = wave_tiles < half ? wave_tiles : half;
            lanes = lanes_get();
            if (lanes == nullptr) return (int)cudaErrorUnknown;
            chunk = lane_tiles;
        }
    }
    // one call at a time enqueues on the shared auxiliary stream / events
    std::unique_lock<std::mutex> lane_lock;
    if (lanes != nullptr) lane_lock = std::unique_lock<std::mutex>(lanes->mu);
    const bool ws_per_tile = prm.tiles_total <= prm.chunk_tiles;
    float2* const zws0 = prm.zws;
    cudaStream_t const st0 = st;
    bool forked = false;
    int nchunk = 0;
    for (int base = 0; base < prm.tiles_total; base += chunk, ++nchunk) {
        prm.tile_base = base;
        prm.chunk_now = prm.tiles_total - base < chunk ? prm.tiles_total - base : chunk;
        if (lanes != nullptr) {
            const int lane = nchunk & 1;
            st = lane ? lanes->aux : st0;
            prm.zws = zws0 + (long long)(ws_per_tile ? base : lane * (prm.chunk_tiles / 2)) * (D * D) * 4096;
            prm.eq = eq0 != nullptr ? eq0 + (long long)(ws_per_tile ? base : lane * (prm.chunk_tiles / 2)) * npp : nullptr;
            if (!forked) {  // fork BEFORE the first chunk is queued: the auxiliary lane waits only for earlier work
                if (cudaError_t e = cudaEventRecord(lanes->fork, st0)) return (int)e;
                if (cudaError_t e = cudaStreamWaitEvent(lanes->aux, lanes->fork, 0)) return (int)e;
                forked = true;
            }
        }
        const int units = prm.chunk_now * npp;
        const int grid_f = units < sms * per_sm_f ? units : sms * per_sm_f;
        const int grid_i = units < sms * per_sm_i ? units : sms * per_sm_i;
        cudaError_t e;
        // ring-fed forward launch (sub_ring.cuh): bit-identical, 32.4 vs 35.7 us alone under ncu, but the STEP is 4 %
        // slower (77.0 vs 74.1 us; rgb 190 vs 153 us) -- its 207 KB CTA cannot become resident while the previous
        // call's inverse CTAs drain, and 72 KB of ring is not enough bytes in flight per SM -- so it is opt-in
        static const bool use_ring = getenv("TFCFFT_SUB_FWD_RING") != nullptr;
        if (lanes == nullptr) prm.eq = eq0;
        if (prm.fine_deps || (D == 4 && use_ring)) prm.eq = nullptr;  // the opt-in schedules do not write the flags
        int ring_rc = TFCFFT_ERR_STRIDE;
        if (D == 4 && use_ring && !prm.fine_deps && ring_addressable<T>(prm)) ring_rc = launch_sub_fwd_ring<T, LUMA3>(prm, st);
        if (ring_rc == 0) {
            e = cudaSuccess;
            g_launches--;  // counted by the ring launcher; the common increment follows below
        } else if (ring_rc != TFCFFT_ERR_STRIDE) {
            return ring_rc;
        } else if (cluster8) {  // 4-CTA clusters
            e = launch_pdl(kf8, grid_f & ~3, SubCfg::NT_FWD, SubCfg::SMEM_FWD, st, prm);
        } else if (cluster) {  // 2-CTA clusters: full-sector loads, halves exchanged through DSMEM
            e = launch_pdl(kf4, grid_f & ~1, SubCfg::NT_FWD, SubCfg::SMEM_FWD, st, prm);
        } else {
            e = launch_pdl(kf, grid_f, SubCfg::NT_FWD, SubCfg::SMEM_FWD, st, prm);
        }
        if (e != cudaSuccess) return (int)e;
        g_launches++;
        e = launch_combine(D, prm.chunk_now * prm.parts, prm, st);
        if (e != cudaSuccess) return (int)e;
        g_launches++;
        if (prm.grad) {
            // one launch order (no lanes): the last inverse launch follows every combine launch -> it carries the
            // CTAs (one cluster's worth) that sum the partial sums
            const bool last = prm.tile_base + prm.chunk_now >= prm.tiles_total;
            // D = 4: one CTA per (tile, row phase) without clusters (measured: profiles/r02_final_ab.txt); the 2-CTA cluster
            // form stays for the tile-granular-dependency schedule and behind TFCFFT_INV4_CLUSTER=1
            static const bool inv4_cluster = getenv("TFCFFT_INV4_CLUSTER") != nullptr;
            const bool wide = cluster && !prm.fine_deps && !inv4_cluster;
            const int fin = (prm.defer_finish && lanes == nullptr && last) ? (cluster8 ? 4 : (cluster && !wide) ? 2 : 1) : 0;
            prm.fin_ctas = fin;
            if (wide) {
                const int rows = prm.chunk_now * 4, cap = sms * per_sm_i4w;
                e = launch_pdl(ki4w, (rows < cap ? rows : cap) + fin, SubCfg::NT_FWD, SubCfg::SMEM_FWD, st, prm);
            } else if (cluster8) {
                e = launch_pdl(ki8, (grid_i & ~3) + fin, SubCfg::NT_INV, SubCfg::SMEM_INV, st, prm);
            } else if (cluster) {
                e = launch_pdl(ki4, (grid_i & ~1) + fin, SubCfg::NT_INV, SubCfg::SMEM_INV, st, prm);
            } else {
                e = launch_pdl(ki, grid_i + fin, SubCfg::NT_INV, SubCfg::SMEM_INV, st, prm);
            }
            prm.fin_ctas = 0;
            if (e != cudaSuccess) return (int)e;
            g_launches++;
        }
    }
    if (forked) {  // join: the caller's stream continues after the auxiliary lane
        if (cudaError_t e = cudaEventRecord(lanes->join, lanes->aux)) return (int)e;
        if (cudaError_t e = cudaStreamWaitEvent(st0, lanes->join, 0)) return (int)e;
    }
    if (prm.defer_finish && lanes != nullptr) {  // both lanes are behind the caller's stream here
        // a plain launch: ordered behind everything the join put in front of it
        sub_finalize_kernel<TFC_DT><<<1, 128, 0, st0>>>(prm);
        if (cudaError_t e = cudaGetLastError()) return (int)e;
        g_launches++;
    }
    return 0;
}

}  // namespace

int TFC_FN(launch_sub)(bool luma3, const Params& prm, cudaStream_t st) {
    return luma3 ? launch_sub<TFC_T, true>(prm, st) : launch_sub<TFC_T, false>(prm, st);
}

}  // namespace tfcfft
