// Store-path throughput of one B200 for the output patterns of the smoothed-IvE apply: 68763 rows x 512 doubles written
// once by 148 CTAs of 16 warps, each warp owning a 32-field (256-byte) strip of the CTA's rows.
//   P0: one STG.128 per lane, a warp instruction = 512 contiguous bytes of one row (two warps' strips)
//   P1: DMMA C-fragment layout: a warp instruction = 8 rows x 64 bytes
//   P2: a warp instruction = 2 rows x 256 bytes
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o store_rate store_rate.cu && ./store_rate
#include <cstdio>
#include <cuda_runtime.h>

template <int P>
__global__ void __launch_bounds__(512) store_kernel(double *Y, int nrow, double v)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r0 = (int)((long long)nrow * blockIdx.x / gridDim.x), r1 = (int)((long long)nrow * (blockIdx.x + 1) / gridDim.x);
    const double2 val = make_double2(v, v + lane);
    for (int rb = r0; rb < r1; rb += 32) {
        if (P == 0) {
            // warp w: rows rb + 2w, 2w + 1; 512 fields = 2 instructions of 512 B per row... use 8 per row
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const int r = rb + 2 * warp + (k >> 3);
                if (r < r1) *reinterpret_cast<double2 *>(Y + (size_t)r * 512 + (k & 7) * 64 + 2 * lane) = val;
            }
        } else if (P == 1) {
            const int g = lane >> 2, q = lane & 3;
#pragma unroll
            for (int t = 0; t < 4; ++t)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int r = rb + 8 * t + g;
                    if (r < r1) *reinterpret_cast<double2 *>(Y + (size_t)r * 512 + 32 * warp + 8 * j + 2 * q) = val;
                }
        } else {
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const int r = rb + 2 * k + (lane >> 4);
                if (r < r1) *reinterpret_cast<double2 *>(Y + (size_t)r * 512 + 32 * warp + 2 * (lane & 15)) = val;
            }
        }
    }
}

int main()
{
    const int nrow = 68763;
    double *Y;
    cudaMalloc(&Y, sizeof(double) * 512 * (size_t)nrow);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int p = 0; p < 3; ++p)
        for (int grid : {148, 296}) {
            float best = 1e9f;
            for (int rep = 0; rep < 5; ++rep) {
                cudaEventRecord(e0);
                for (int k = 0; k < 10; ++k) {
                    if (p == 0) store_kernel<0><<<grid, 512>>>(Y, nrow, 1.0 + k);
                    else if (p == 1) store_kernel<1><<<grid, 512>>>(Y, nrow, 1.0 + k);
                    else store_kernel<2><<<grid, 512>>>(Y, nrow, 1.0 + k);
                }
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms;
                cudaEventElapsedTime(&ms, e0, e1);
                if (ms / 10 < best) best = ms / 10;
            }
            printf("P%d grid %d: %.1f us  %.0f GB/s\n", p, grid, best * 1e3, 8.0 * 512 * nrow / best / 1e6);
        }
    printf("cuda status: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
