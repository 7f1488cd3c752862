// k_line.cu -- 64 x 64 tiles (16-patch loss at 256 x 256: the north-star shape).
//   line_kernel      thread-per-line kernel, one 64-thread CTA per tile, blocking loads (line_tile.cuh); kept as the
//                    A/B baseline (TFCFFT_LINE_V1=1) and for inputs the bulk-copy engine cannot address
//   line_ring_kernel the same transforms fed by an asynchronous shared-memory ring: one producer warp streams raw
//                    row slabs with cp.async.bulk + mbarrier completion, worker groups convert and transform
//                    (line_ring.cuh)
#include "launchers.h"
#include "line_tile.cuh"
#include "line_ring.cuh"

namespace tfcfft {

// Thread-per-line kernel for 64 x 64 tiles (line_tile.cuh): 64 threads = one tile, six CTAs per SM.
template <typename T, bool LUMA3>
__global__ void __launch_bounds__(LineCfg::NT, 6) line_kernel(const __grid_constant__ Params prm) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float2* s = reinterpret_cast<float2*>(smem_raw);
    BlockCtxT<LineCfg::NT> ctx{(int)threadIdx.x, nullptr};
    pdl_wait();
    int iter = 0;
    for (int tile = blockIdx.x; tile < prm.tiles_total; tile += gridDim.x, ++iter) {
        float a = 0.f, p = 0.f;
        ctx.trace = (prm.trace != nullptr && iter < 6) ? prm.trace + ((long long)blockIdx.x * 6 + iter) * 16 : nullptr;
        const int nt = tile + (int)gridDim.x;
        line_process<T, LUMA3>(ctx, prm, tile, s, a, p, nt < prm.tiles_total ? nt : -1, false);
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

namespace {

template <typename T, bool LUMA3>
int launch_line_v1(const Params& prm, cudaStream_t st) {
    auto kernel = line_kernel<T, LUMA3>;
    constexpr size_t smem = LineCfg::SMEM;
    static KernelFacts facts;
    int per_sm = 1;
    if (int rc = facts.get(kernel, LineCfg::NT, smem, &per_sm)) return rc;
    const long long cap = (long long)device_sms() * per_sm;
    const int grid = (int)(prm.tiles_total < cap ? prm.tiles_total : cap);
    if (cudaError_t e2 = launch_pdl(kernel, grid, LineCfg::NT, smem, st, prm)) return (int)e2;
    g_launches++;
    return 0;
}

template <typename T, bool LUMA3>
int launch_line(const Params& prm, cudaStream_t st) {
    static const bool v1 = getenv("TFCFFT_LINE_V1") != nullptr;
    if (!v1 && ring_addressable<T>(prm)) {
        const int rc = launch_line_ring<T, LUMA3>(prm, st);
        if (rc != TFCFFT_ERR_STRIDE) return rc;  // no tensor map for these inputs: blocking-load kernel below
    }
    return launch_line_v1<T, LUMA3>(prm, st);
}

}  // namespace

int TFC_FN(launch_line)(bool luma3, const Params& prm, cudaStream_t st) {
    return luma3 ? launch_line<TFC_T, true>(prm, st) : launch_line<TFC_T, false>(prm, st);
}

}  // namespace tfcfft
