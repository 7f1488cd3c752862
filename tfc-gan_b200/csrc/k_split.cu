// k_split.cu -- split kernels: row and column transforms in separate launches that exchange the complex spectrum
// through an L2-sized workspace chunk.  Used for the spectra-materialising modes at P >= 256 and when forced
// (TFCFFT_FORCE_SPLIT); the loss path at P = 128 / 256 / 512 runs on the sub-tile pipeline (k_sub.cu).
//   split_rows_fwd / split_cols / split_rows_inv   (spectral_core.cuh)
#include "launchers.h"

namespace tfcfft {

template <int P> struct SplitCfg {
    static constexpr int NT = 256;
    static constexpr size_t SMEM_ROWS = ((size_t)Split<P>::RS * (P + 1) + P) * sizeof(float2);
    static constexpr size_t SMEM_COLS = ((size_t)P * (2 * Split<P>::GS + 1) + P) * sizeof(float2);
};

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

#if TFC_DT == 0
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
#endif

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

namespace {

template <int P, typename T, bool LUMA3>
int launch_split(Params prm, cudaStream_t st) {
    using Sp = Split<P>;
    auto k1 = split_rows_fwd_kernel<P, T, LUMA3>;
    auto k3 = split_rows_inv_kernel<P, T, LUMA3>;
    static KernelFacts f1, f3;
    if (int rc = f1.get(k1, SplitCfg<P>::NT, SplitCfg<P>::SMEM_ROWS, nullptr)) return rc;
    if (int rc = split_cols_facts(P)) return rc;
    if (int rc = f3.get(k3, SplitCfg<P>::NT, SplitCfg<P>::SMEM_ROWS, nullptr)) return rc;
    for (int base = 0; base < prm.tiles_total; base += prm.chunk_tiles) {
        prm.tile_base = base;
        const int nt = prm.tiles_total - base < prm.chunk_tiles ? prm.tiles_total - base : prm.chunk_tiles;
        if (cudaError_t e = launch_pdl(k1, nt * Sp::ROW_SLABS, SplitCfg<P>::NT, SplitCfg<P>::SMEM_ROWS, st, prm)) return (int)e;
        g_launches++;
        if (cudaError_t e = launch_split_cols(P, nt * Sp::PARTS, prm, st)) return (int)e;
        g_launches++;
        if (prm.grad) {
            if (cudaError_t e = launch_pdl(k3, nt * Sp::ROW_SLABS, SplitCfg<P>::NT, SplitCfg<P>::SMEM_ROWS, st, prm)) return (int)e;
            g_launches++;
        }
    }
    return 0;
}

}  // namespace

int TFC_FN(launch_split)(int p, bool luma3, const Params& prm, cudaStream_t st) {
#define TFC_SPL(P) (luma3 ? launch_split<P, TFC_T, true>(prm, st) : launch_split<P, TFC_T, false>(prm, st))
    switch (p) {
        case 64: return TFC_SPL(64);
        case 128: return TFC_SPL(128);
        case 256: return TFC_SPL(256);
        case 512: return TFC_SPL(512);
    }
#undef TFC_SPL
    return TFCFFT_ERR_SHAPE;
}

#if TFC_DT == 0
// the column launch does not depend on the element type: one copy, in the fp32 object
namespace {
template <int P>
int cols_facts() {
    static KernelFacts f;
    return f.get(split_cols_kernel<P>, SplitCfg<P>::NT, SplitCfg<P>::SMEM_COLS, nullptr);
}
}  // namespace
int split_cols_facts(int p) {
    switch (p) {
        case 64: return cols_facts<64>();
        case 128: return cols_facts<128>();
        case 256: return cols_facts<256>();
        case 512: return cols_facts<512>();
    }
    return TFCFFT_ERR_SHAPE;
}
cudaError_t launch_split_cols(int p, int grid, const Params& prm, cudaStream_t st) {
    switch (p) {
        case 64: return launch_pdl(split_cols_kernel<64>, grid, SplitCfg<64>::NT, SplitCfg<64>::SMEM_COLS, st, prm);
        case 128: return launch_pdl(split_cols_kernel<128>, grid, SplitCfg<128>::NT, SplitCfg<128>::SMEM_COLS, st, prm);
        case 256: return launch_pdl(split_cols_kernel<256>, grid, SplitCfg<256>::NT, SplitCfg<256>::SMEM_COLS, st, prm);
        case 512: return launch_pdl(split_cols_kernel<512>, grid, SplitCfg<512>::NT, SplitCfg<512>::SMEM_COLS, st, prm);
    }
    return cudaErrorInvalidValue;
}
#endif

}  // namespace tfcfft
