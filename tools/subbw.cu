// subbw.cu -- load-phase ceiling of the forward sub-image launch (global 256 x 256 loss, batch 64, D = 4).
// mode 0: one CTA per sub-image PAIR, 8-byte loads that use half of every 32-byte sector (the other half goes to
//         the CTA of the neighbouring pair): every sector crosses the L2 -> SM fabric twice.
// mode 1: CTA pairs share a row phase; each CTA reads half of the rows with full 16-byte loads (what a 2-CTA
//         cluster exchanging halves through distributed shared memory would load).
// Both read the same 100.7 MB once from HBM.  Not part of the product; results go to profiles/.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(128, 3) sub_pattern(const float* __restrict__ fake, const float* __restrict__ real,
                                                      float* __restrict__ sink, int units) {
    float acc = 0.f;
    for (int u = blockIdx.x; u < units; u += gridDim.x) {
        const int tile = u >> 3, p = (u >> 1) & 3, i = u & 1;
        const size_t base = (size_t)tile * 3 * 65536;
        if (MODE == 0) {
            for (int it0 = threadIdx.x; it0 < 4096; it0 += 4 * 128) {
                float2 v[4][6];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int it = it0 + k * 128, b = it & 63, a = it >> 6, x = 4 * b + 2 * i, y = 4 * a + p;
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        v[k][c] = *reinterpret_cast<const float2*>(fake + base + (size_t)c * 65536 + y * 256 + x);
                        v[k][3 + c] = *reinterpret_cast<const float2*>(real + base + (size_t)c * 65536 + y * 256 + x);
                    }
                }
#pragma unroll
                for (int k = 0; k < 4; ++k)
#pragma unroll
                    for (int c = 0; c < 6; ++c) acc += v[k][c].x + v[k][c].y;
            }
        } else {
            for (int it0 = threadIdx.x; it0 < 2048; it0 += 4 * 128) {
                float4 v[4][6];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int it = it0 + k * 128, b = it & 63, a = (it >> 6) + 32 * i, x = 4 * b, y = 4 * a + p;
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        v[k][c] = *reinterpret_cast<const float4*>(fake + base + (size_t)c * 65536 + y * 256 + x);
                        v[k][3 + c] = *reinterpret_cast<const float4*>(real + base + (size_t)c * 65536 + y * 256 + x);
                    }
                }
#pragma unroll
                for (int k = 0; k < 4; ++k)
#pragma unroll
                    for (int c = 0; c < 6; ++c) acc += v[k][c].x + v[k][c].y + v[k][c].z + v[k][c].w;
            }
        }
    }
    if (acc == 12345.678f) sink[0] = acc;
}

int main() {
    const int N = 64, units = N * 8, POOL = 4;
    const size_t elems = (size_t)N * 3 * 65536;
    float *fake[POOL], *real[POOL], *sink;
    for (int i = 0; i < POOL; ++i) {
        cudaMalloc(&fake[i], elems * 4);
        cudaMalloc(&real[i], elems * 4);
        cudaMemset(fake[i], 0, elems * 4);
        cudaMemset(real[i], 0, elems * 4);
    }
    cudaMalloc(&sink, 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const double bytes = 2.0 * elems * 4;
    // lone-CTA condition (second-round units of the real kernel): one unit per CTA on few CTAs
    for (int g : {68, 148, 296})
        for (int mode = 0; mode < 2; ++mode) {
            cudaEventRecord(e0);
            for (int it = 0; it < 40; ++it) {
                if (mode == 0) sub_pattern<0><<<g, 128>>>(fake[it % POOL], real[it % POOL], sink, g);
                else sub_pattern<1><<<g, 128>>>(fake[it % POOL], real[it % POOL], sink, g);
            }
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            printf("grid %d mode %d: %.1f us per launch (one 196 KB unit per CTA)\n", g, mode, ms * 1e3 / 40);
        }
    for (int mode = 0; mode < 2; ++mode)
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            for (int it = 0; it < 40; ++it) {
                if (mode == 0) sub_pattern<0><<<444, 128>>>(fake[it % POOL], real[it % POOL], sink, units);
                else sub_pattern<1><<<444, 128>>>(fake[it % POOL], real[it % POOL], sink, units);
            }
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            printf("mode %d (%s): %.1f us per launch, %.0f GB/s of %.1f MB\n", mode, mode ? "16-byte loads, half the rows" : "8-byte half-sector loads",
                   ms * 1e3 / 40, bytes * 40 / (ms * 1e-3) / 1e9, bytes / 1e6);
        }
    printf("status: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
