// util.cuh -- small kernels shared by build.cu and smooth.cu (static: one copy per translation unit).
#pragma once
#include "common.cuh"

namespace ib {

// ptr[r] = first j with idx[j] >= r  (idx ascending); ptr[nseg] = n
static __global__ void seg_ptr_kernel(const int32_t *idx, int64_t n, int64_t nseg, int32_t *ptr)
{
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    int32_t r = idx[j];
    int32_t prev = j > 0 ? idx[j - 1] : -1;
    for (int32_t q = prev + 1; q <= r; ++q) ptr[q] = (int32_t)j;
    if (j == n - 1)
        for (int64_t q = (int64_t)r + 1; q <= nseg; ++q) ptr[q] = (int32_t)n;
}

static __global__ void iota_u32_kernel(uint32_t *p, int64_t n)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = (uint32_t)i;
}
static inline int bits_for(int64_t n)
{
    int b = 1;
    while ((int64_t(1) << b) < n) ++b;
    return b;
}


}   // namespace ib
