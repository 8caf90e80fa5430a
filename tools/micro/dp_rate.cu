// Measures fp64 throughput of one B200 with independent DFMA / DMUL+DADD chains and the shared-memory read rate
// (LDS.128), to bound kernels whose arithmetic intensity is not negligible (the smoothed IvE apply).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dp_rate dp_rate.cu && ./dp_rate
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void dp_kernel(double *out, double a, double b, int iters)
{
    double x[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = threadIdx.x * 1e-3 + k;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (MODE == 0) x[k] = fma(x[k], a, b);
            else x[k] = __dadd_rn(__dmul_rn(x[k], a), b);
        }
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void lds_kernel(double *out, int iters)
{
    extern __shared__ double2 sm[];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = make_double2(i, 1.0);
    __syncthreads();
    double2 acc = make_double2(0, 0);
    int idx = threadIdx.x;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            double2 v = sm[(idx + k * 256) & 4095];
            acc.x += v.x; acc.y += v.y;
        }
        idx = (idx + 37) & 4095;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y;
}

// fp64 tensor-core rate: mma.sync.m8n8k4.f64, NT independent accumulator tiles per warp
template <int NT>
__global__ void dmma_kernel(double *out, double a, double b, int iters)
{
    double c0[NT], c1[NT];
#pragma unroll
    for (int k = 0; k < NT; ++k) { c0[k] = threadIdx.x * 1e-3 + k; c1[k] = k; }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < NT; ++k)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0[k]), "+d"(c1[k]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < NT; ++k) s += c0[k] + c1[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main()
{
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double *out;
    cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int mode = 0; mode < 2; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) dp_kernel<0><<<sms * 8, 256>>>(out, 1.0000001, 1e-9, iters);
            else dp_kernel<1><<<sms * 8, 256>>>(out, 1.0000001, 1e-9, iters);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            double ops = (double)sms * 8 * 256 * 8 * iters * (mode == 0 ? 1 : 2);      // DP instructions
            if (rep) printf("%s: %.3f ms, %.2f T DP instr/s (%.1f per clk per SM at 1.9 GHz), %.2f TFLOP/s\n", mode == 0 ? "DFMA" : "DMUL+DADD", ms,
                            ops / ms / 1e9, ops / ms / 1e9 * 1e12 / sms / 1.9e9 / 1e12 * 1.0, (mode == 0 ? 2 : 1) * ops / ms / 1e9);
        }
    }
    for (int nt = 0; nt < 2; ++nt)
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (nt == 0) dmma_kernel<8><<<sms * 8, 256>>>(out, 1.0000001, 1e-9, iters / 4);
            else dmma_kernel<16><<<sms * 4, 256>>>(out, 1.0000001, 1e-9, iters / 4);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            double mmas = (double)sms * (nt == 0 ? 8 : 4) * 8 * (nt == 0 ? 8 : 16) * (iters / 4);      // warp-level instructions
            if (rep) printf("DMMA m8n8k4 (%d tiles/warp): %.3f ms, %.2f G mma/s, %.2f TFLOP/s (%.2f mma per clk per SM at 1.9 GHz)\n", nt == 0 ? 8 : 16, ms,
                            mmas / ms / 1e6, mmas * 512 / ms / 1e9, mmas / ms / 1e6 * 1e9 / sms / 1.9e9);
        }
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        lds_kernel<<<sms * 4, 256, 65536>>>(out, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        double bytes = (double)sms * 4 * 256 * 8 * iters * 16;
        if (rep) printf("LDS.128: %.3f ms, %.2f TB/s (%.1f B per clk per SM at 1.9 GHz)\n", ms, bytes / ms / 1e9, bytes / ms / 1e9 * 1e12 / sms / 1.9e9 / 1e12);
    }
    printf("cuda status: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
