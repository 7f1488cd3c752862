// k_resident.cu -- generic resident-tile kernel (P <= 128, all modes) and the packed tile-pair kernel (P = 64, A/B path).
//
//   resident_kernel      persistent CTAs, one complex tile resident in shared memory:
//                        load+luma+pack -> row FFT -> column FFT -> loss + spectral gradient ->
//                        inverse column FFT -> inverse row FFT -> gradient store.
//   pair_kernel          two tiles per CTA as packed f32x2 lanes, warp-specialised loader (pair_tile.cuh)
#include "launchers.h"
#include "pair_tile.cuh"

namespace tfcfft {

template <int P> struct ResidentCfg {
    static constexpr int NT = P <= 16 ? 64 : P <= 32 ? 128 : P <= 64 ? 256 : 512;
    static constexpr size_t SMEM = ((size_t)P * (P + 1) + P) * sizeof(float2);
};

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
//                           spectral gradient, inverse FFTs;
//   warps 16..23 (load):    stream pair i+1 from HBM (coalesced 128-bit loads), fold luma, pack (A, B) and
//                           fill the other work buffer, and write the finished gradient tiles of pair i-1 out.
// Hand-off with named barriers: FULL[b] (loaders arrive, compute waits), DONE[b] (compute arrives when
// buffer b holds the gradient tiles, loaders wait, store them and refill the buffer).
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

namespace {

template <int P, typename T, bool LUMA3>
int launch_resident(const Params& prm, cudaStream_t st) {
    auto kernel = resident_kernel<P, T, LUMA3>;
    constexpr size_t smem = ResidentCfg<P>::SMEM;
    constexpr int nt = ResidentCfg<P>::NT;
    static KernelFacts facts;
    int per_sm = 1;
    if (int rc = facts.get(kernel, nt, smem, &per_sm)) return rc;
    const long long cap = (long long)device_sms() * per_sm;  // persistent: one wave
    const int grid = (int)(prm.tiles_total < cap ? prm.tiles_total : cap);
    kernel<<<grid, nt, smem, st>>>(prm);
    g_launches++;
    TFC_LAUNCH_CHECK();
    return 0;
}

template <typename T, bool LUMA3>
int launch_pair(const Params& prm, cudaStream_t st) {
    constexpr int P = 64;
    auto kernel = pair_kernel<P, T, LUMA3>;
    constexpr size_t smem = PairCfg<P>::SMEM;
    constexpr int nt = PairCfg<P>::NT;
    static KernelFacts facts;
    if (int rc = facts.get(kernel, nt, smem, nullptr)) return rc;
    const long long npairs = ((long long)prm.tiles_total + 1) / 2;
    const long long cap = device_sms();  // persistent, warp-specialised: one CTA per SM
    const int grid = (int)(npairs < cap ? npairs : cap);
    kernel<<<grid, nt, smem, st>>>(prm);
    g_launches++;
    TFC_LAUNCH_CHECK();
    return 0;
}

}  // namespace

int TFC_FN(launch_resident)(int p, bool luma3, const Params& prm, cudaStream_t st) {
#define TFC_RES(P) (luma3 ? launch_resident<P, TFC_T, true>(prm, st) : launch_resident<P, TFC_T, false>(prm, st))
    switch (p) {
        case 16: return TFC_RES(16);
        case 32: return TFC_RES(32);
        case 64: return TFC_RES(64);
        case 128: return TFC_RES(128);
    }
#undef TFC_RES
    return TFCFFT_ERR_SHAPE;
}

int TFC_FN(launch_pair)(bool luma3, const Params& prm, cudaStream_t st) {
    return luma3 ? launch_pair<TFC_T, true>(prm, st) : launch_pair<TFC_T, false>(prm, st);
}

}  // namespace tfcfft
