// tilebw.cu -- memory-system ceiling for the 64x64-tile access pattern of the 16-patch loss.
// Each 64-thread CTA loads one tile of fake and real (3 planes each, 64 rows x 256 B at 1 KB pitch) with the
// line kernel's exact load pattern and writes a 3-plane gradient tile; no transforms.  Compared with a plain
// streaming kernel moving the same bytes.  Not part of the product; results go to profiles/.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(64, 6) tile_pattern(const float* __restrict__ fake, const float* __restrict__ real,
                                                      float* __restrict__ grad, int tiles, int ni_sel) {
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int n = tile >> 4, py = (tile >> 2) & 3, px = tile & 3;
        const size_t base = (size_t)n * 3 * 65536 + (size_t)py * 64 * 256 + px * 64;
        float4 acc = make_float4(0, 0, 0, 0);
        for (int it0 = threadIdx.x; it0 < 1024; it0 += 4 * 64) {
            float4 v[4][6];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int it = it0 + u * 64, x = (it & 15) * 4, y = it >> 4;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    v[u][c] = *reinterpret_cast<const float4*>(fake + base + (size_t)c * 65536 + y * 256 + x);
                    v[u][3 + c] = *reinterpret_cast<const float4*>(real + base + (size_t)c * 65536 + y * 256 + x);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int c = 0; c < 6; ++c) { acc.x += v[u][c].x; acc.y += v[u][c].y; acc.z += v[u][c].z; acc.w += v[u][c].w; }
        }
        for (int it = threadIdx.x; it < 1024; it += 64) {
            const int x = (it & 15) * 4, y = it >> 4;
#pragma unroll
            for (int c = 0; c < 3; ++c) *reinterpret_cast<float4*>(grad + base + (size_t)c * 65536 + y * 256 + x) = acc;
        }
    }
}

__global__ void __launch_bounds__(256) stream_pattern(const float4* __restrict__ fake, const float4* __restrict__ real,
                                                      float4* __restrict__ grad, size_t n4) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const float4 a = fake[i], b = real[i];
        grad[i] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
    }
}

int main() {
    const int N = 256, tiles = N * 16;
    const size_t elems = (size_t)N * 3 * 65536;
    float *fake[2], *real[2], *grad;
    for (int i = 0; i < 2; ++i) {
        cudaMalloc(&fake[i], elems * 4);
        cudaMalloc(&real[i], elems * 4);
        cudaMemset(fake[i], 0, elems * 4);
        cudaMemset(real[i], 0, elems * 4);
    }
    cudaMalloc(&grad, elems * 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const double bytes = 3.0 * elems * 4;
    for (int mode = 0; mode < 2; ++mode) {
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            for (int it = 0; it < 20; ++it) {
                if (mode == 0) tile_pattern<<<148 * 6, 64>>>(fake[it & 1], real[it & 1], grad, tiles, 4);
                else stream_pattern<<<148 * 8, 256>>>((const float4*)fake[it & 1], (const float4*)real[it & 1], (float4*)grad, elems / 4);
            }
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            printf("%s: %.1f us per pass, %.0f GB/s (read 2 + write 1 tensors of %.0f MB)\n", mode ? "stream" : "tiles ", ms * 1e3 / 20,
                   bytes * 20 / (ms * 1e-3) / 1e9, elems * 4 / 1e6);
        }
    }
    printf("status: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
