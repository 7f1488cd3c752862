// kernel_common.cuh -- device helpers shared by the kernel translation units (block reductions, the last-CTA
// finalise, named barriers, programmatic dependent launch) and the host helpers every launcher uses (per-device
// caches of SM count / occupancy / function attributes, the launch counter).
//
// The library is built from several translation units (k_*.cu + tfcfft_api.cu) so that the build parallelises;
// launchers.h declares the host entry of each unit.
#pragma once
#include <atomic>
#include <cstdlib>
#include <mutex>

#include "spectral_core.cuh"

namespace tfcfft {

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
extern std::atomic<long long> g_launches;  // tfcfft_api.cu

constexpr int kMaxDevices = 64;

inline int current_device() {
    int dev = 0;
    cudaGetDevice(&dev);
    return dev < 0 || dev >= kMaxDevices ? 0 : dev;
}

// SM count, cached per device (one attribute query per device per process)
inline int device_sms() {
    static std::atomic<int> cache[kMaxDevices];
    const int dev = current_device();
    int v = cache[dev].load(std::memory_order_relaxed);
    if (v == 0) {
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        if (v < 1) v = 1;
        cache[dev].store(v, std::memory_order_relaxed);
    }
    return v;
}

// Per-(kernel instantiation, device) launch facts: the opt-in shared-memory attribute is set once and the
// occupancy is queried once; later calls are two relaxed atomic loads.  One static KernelFacts per launcher
// template instantiation.
struct KernelFacts {
    std::atomic<int> ready[kMaxDevices];
    std::atomic<int> per_sm[kMaxDevices];
    std::mutex mu;
    // returns 0 and the cached occupancy, or a cudaError_t
    template <class K>
    int get(K kernel, int threads, size_t smem, int* occ) {
        const int dev = current_device();
        if (!ready[dev].load(std::memory_order_acquire)) {
            std::lock_guard<std::mutex> lk(mu);
            if (!ready[dev].load(std::memory_order_relaxed)) {
                if (smem > 48 * 1024) {
                    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                    if (e != cudaSuccess) return (int)e;
                }
                int o = 0;
                cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kernel, threads, smem);
                if (e != cudaSuccess) return (int)e;
                per_sm[dev].store(o < 1 ? 1 : o, std::memory_order_relaxed);
                ready[dev].store(1, std::memory_order_release);
            }
        }
        if (occ) *occ = per_sm[dev].load(std::memory_order_relaxed);
        return 0;
    }
};

#define TFC_LAUNCH_CHECK()                       \
    do {                                         \
        cudaError_t e__ = cudaGetLastError();    \
        if (e__ != cudaSuccess) return (int)e__; \
    } while (0)

// launch with the programmatic-stream-serialization attribute (pdl_wait / pdl_release below)
template <class K, class P>
cudaError_t launch_pdl(K kernel, int grid, int block, size_t smem, cudaStream_t st, const P& prm, int cluster = 0) {
    static const bool off = getenv("TFCFFT_NO_PDL") != nullptr;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (!off) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    if (cluster > 1) {
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = (unsigned)cluster;
        attr[na].val.clusterDim.y = 1;
        attr[na].val.clusterDim.z = 1;
        ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    return cudaLaunchKernelEx(&cfg, kernel, prm);
}
// the same for kernels with a plain argument list (no cluster)
template <class... KA, class... A>
cudaError_t launch_pdl_args(void (*kernel)(KA...), int grid, int block, cudaStream_t st, A... args) {
    static const bool off = getenv("TFCFFT_NO_PDL") != nullptr;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)block);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = off ? 0 : 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KA>(args)...);
}

#ifdef __CUDACC__
// ---------------------------------------------------------------------------------------------------------------
// device side
// ---------------------------------------------------------------------------------------------------------------
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

// Fixed-order sum of all partials in double by ONE block; writes the outputs.  Every partial must be visible.
__device__ __forceinline__ void finalize_sums(const Params& prm) {
    __shared__ double dred[2][32];
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
    }
}

// Last block to arrive sums all partials in a fixed order (double) and writes the outputs;
// the ticket counter is left at zero for the next call.
__device__ __forceinline__ void finish(const Params& prm, unsigned total_blocks) {
    __shared__ bool last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        last = (atomicAdd(prm.counter, 1u) == total_blocks - 1);
    }
    __syncthreads();
    if (!last) return;
    finalize_sums(prm);
    if (threadIdx.x == 0) *prm.counter = 0u;
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

// Programmatic dependent launch (sm_90+): consecutive launches of this library are chained with the
// programmatic-stream-serialization attribute, so the next grid is scheduled while the previous one drains and its
// CTAs sit in `griddepcontrol.wait` until that grid has completed and flushed -- the launch latency between
// dependent kernels (~2 us each, a few percent of a 100 us step) overlaps the tail.  Both instructions are no-ops in
// a launch without the attribute.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// Issued by every CTA when its own work is done: releasing earlier lets the dependent grid's CTAs take SM slots
// that this grid's not-yet-started CTAs need (measured: 118 us instead of 98 us per step).
__device__ __forceinline__ void pdl_release() { asm volatile("griddepcontrol.launch_dependents;"); }
#endif  // __CUDACC__

}  // namespace tfcfft
