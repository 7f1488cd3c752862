// ubench.cu -- issue-rate microbenchmarks that size the FFT kernels (B200, sm_100a).
// Prints per-SM throughput (thread-ops per clock) of scalar vs packed-f32x2 arithmetic, MUFU,
// atan2f, and shared-memory 64/128-bit access.  Not part of the product; results go to profiles/.
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>

#define ITERS 2048
#define ILP 8

template <int OP>
__global__ void __launch_bounds__(1024) k_arith(float* out, long long* cyc, float seed) {
    float2 a[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] = make_float2(seed + i + threadIdx.x, seed * 0.5f + i);
    const float2 m = make_float2(1.0000001f, 0.9999999f), c = make_float2(1e-7f, -1e-7f);
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (OP == 0) { a[i].x = a[i].x + c.x; a[i].y = a[i].y + c.y; }                    // 2 FADD
            if (OP == 1) { a[i].x = fmaf(a[i].x, m.x, c.x); a[i].y = fmaf(a[i].y, m.y, c.y); } // 2 FFMA (3 reg)
            if (OP == 2) a[i] = __fadd2_rn(a[i], c);                                           // FADD2
            if (OP == 3) a[i] = __ffma2_rn(a[i], m, c);                                        // FFMA2
            if (OP == 4) a[i] = __fmul2_rn(a[i], m);                                           // FMUL2
            if (OP == 5) { a[i].x = rsqrtf(a[i].x); a[i].y = rsqrtf(a[i].y); }                 // 2 MUFU.RSQ (+fixup)
            if (OP == 6) { a[i].x = atan2f(a[i].y, a[i].x); }                                  // atan2f
            if (OP == 7) { a[i].x = a[i].x * m.x; a[i].y = a[i].y * m.y; }                     // 2 FMUL
            if (OP == 8) { a[i].x = __frcp_rn(a[i].x); }                                       // rcp
            if (OP == 9) { a[i].x = sqrtf(a[i].x); }                                           // sqrt
        }
    }
    const long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += a[i].x + a[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int W>  // W = 8 (LDS.64/STS.64) or 16 (.128)
__global__ void __launch_bounds__(1024) k_smem(float* out, long long* cyc) {
    extern __shared__ float4 sm4[];
    char* base = reinterpret_cast<char*>(sm4);
    for (int i = threadIdx.x; i < 32768 / 16; i += blockDim.x) sm4[i] = make_float4(i, 1, 2, 3);
    __syncthreads();
    float acc = 0;
    const long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            const int off = ((threadIdx.x + i * 1024 + it * 32) * W) & 32767;
            if (W == 8) { float2 v = *reinterpret_cast<float2*>(base + off); acc += v.x; }
            else { float4 v = *reinterpret_cast<float4*>(base + off); acc += v.x; }
        }
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int blocks = sms * 2, nt = 1024;  // 2 x 1024 threads per SM = full occupancy
    float* out;
    long long* cyc;
    cudaMalloc(&out, sizeof(float) * blocks * nt);
    cudaMalloc(&cyc, sizeof(long long) * blocks);
    std::vector<long long> h(blocks);
    const char* names[] = {"FADD x2 (scalar)", "FFMA x2 (scalar)", "FADD2 (f32x2)", "FFMA2 (f32x2)", "FMUL2 (f32x2)",
                           "rsqrtf x2", "atan2f", "FMUL x2 (scalar)", "__frcp_rn", "sqrtf"};
    const double lanes_per[] = {2, 2, 2, 2, 2, 2, 1, 2, 1, 1};
    printf("SMs=%d  (2 blocks x 1024 threads per SM)\n", sms);
#define RUN(OP)                                                                                    \
    {                                                                                              \
        k_arith<OP><<<blocks, nt>>>(out, cyc, 1.5f);                                               \
        k_arith<OP><<<blocks, nt>>>(out, cyc, 1.5f);                                               \
        cudaDeviceSynchronize();                                                                   \
        cudaMemcpy(h.data(), cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);             \
        double avg = 0;                                                                            \
        for (auto v : h) avg += v;                                                                 \
        avg /= blocks;                                                                             \
        const double ops = 2.0 * nt * (double)ITERS * ILP * lanes_per[OP];                         \
        printf("%-20s %8.1f fp32-lane-ops/clk/SM   (%.0f cycles)\n", names[OP], ops / avg, avg);   \
    }
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9)
    for (int w = 8; w <= 16; w += 8) {
        if (w == 8) { k_smem<8><<<blocks, nt, 32768>>>(out, cyc); k_smem<8><<<blocks, nt, 32768>>>(out, cyc); }
        else { k_smem<16><<<blocks, nt, 32768>>>(out, cyc); k_smem<16><<<blocks, nt, 32768>>>(out, cyc); }
        cudaDeviceSynchronize();
        cudaMemcpy(h.data(), cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
        double avg = 0;
        for (auto v : h) avg += v;
        avg /= blocks;
        printf("LDS.%-3d             %8.1f bytes/clk/SM\n", w * 8, 2.0 * nt * (double)ITERS * ILP * w / avg);
    }
    cudaError_t e = cudaGetLastError();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
