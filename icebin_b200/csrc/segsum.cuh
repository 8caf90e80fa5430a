// segsum.cuh -- reduce-by-key over a sorted triplet stream and deterministic segmented sums (shared by build.cu and
// couple.cu; static: one copy per translation unit).
#pragma once
#include "common.cuh"
#include "scan.cuh"

namespace ib {

// ------------------------------------------------------------------------------------------------
// K2: reduce-by-key over the sorted stream (runs summed sequentially, in emission order)
template <class K>
struct HeadLoad {
    const K *keys;
    __device__ uint32_t operator()(int64_t i) const { return (i == 0 || keys[i] != keys[i - 1]) ? 1u : 0u; }
};
struct ReduceApply {
    const uint64_t *keys;
    int cbits;
    int32_t *row, *col, *segstart;
    __device__ void operator()(int64_t i, uint32_t prefix, uint32_t head) const
    {
        if (!head) return;
        const uint64_t k = keys[i];
        row[prefix] = (int32_t)(k >> cbits);
        col[prefix] = (int32_t)(k & ((uint64_t(1) << cbits) - 1));
        segstart[prefix] = (int32_t)i;
    }
};

// Segmented sums out[g] = sum_{j in [ptr[g], ptr[g+1])} val[perm ? perm[j] : j].
// Segments of <= SEG_SEQ elements are summed sequentially by one thread, in storage order -- the order in which Eigen
// sums duplicate triplets and ibmisc's sum() walks a matrix, so those results are bit-identical to the oracle.
// Longer segments (EvA/AvE: thousands of exchange cells per (E, A) pair; A-grid weights at 1 km) would serialise one
// thread for milliseconds; they are queued and reduced by one CTA each: thread-strided partial sums followed by a
// fixed butterfly and a fixed order over the warps, i.e. deterministic, differing from the sequential order only in
// association (<= 1e-13 relative).
constexpr int SEG_SEQ = 64;
static __global__ void seg_sum_short_kernel(const int32_t *__restrict__ ptr, const uint32_t *__restrict__ perm, const double *__restrict__ val,
                                     int64_t nseg, double *__restrict__ out, int32_t *__restrict__ long_list, int *__restrict__ nlong)
{
    int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= nseg) return;
    const int32_t b = ptr[g], e = ptr[g + 1];
    if (e - b > SEG_SEQ) { long_list[atomicAdd(nlong, 1)] = (int32_t)g; return; }
    double s = 0;
    if (perm) for (int32_t j = b; j < e; ++j) s += val[perm[j]];
    else for (int32_t j = b; j < e; ++j) s += val[j];
    out[g] = s;
}
// one CTA per long segment (the A-grid rows at 1 km hold ~17 k entries each, EvA / AvE thousands of duplicates per entry):
// thread-strided partial sums, then a fixed butterfly inside each warp and a fixed order over the warps
static __global__ void __launch_bounds__(256) seg_sum_long_kernel(const int32_t *__restrict__ ptr, const uint32_t *__restrict__ perm,
                                                           const double *__restrict__ val, double *__restrict__ out,
                                                           const int32_t *__restrict__ long_list, const int *__restrict__ nlong)
{
    __shared__ double wsum[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n = *nlong;
    for (int w = blockIdx.x; w < n; w += gridDim.x) {
        const int32_t g = long_list[w];
        const int32_t b = ptr[g], e = ptr[g + 1];
        double s = 0;
        if (perm) for (int32_t j = b + threadIdx.x; j < e; j += 256) s += val[perm[j]];
        else for (int32_t j = b + threadIdx.x; j < e; j += 256) s += val[j];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
        if (lane == 0) wsum[warp] = s;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = wsum[0];
#pragma unroll
            for (int k = 1; k < 8; ++k) t += wsum[k];
            out[g] = t;
        }
        __syncthreads();
    }
}
static void seg_sum(const int32_t *ptr, const uint32_t *perm, const double *val, int64_t nseg, double *out, cudaStream_t s)
{
    if (nseg <= 0) return;
    DBuf<int32_t> long_list((size_t)nseg, s);
    DBuf<int> nlong(1, s);
    nlong.zero();
    seg_sum_short_kernel<<<(unsigned)cdiv(nseg, 128), 128, 0, s>>>(ptr, perm, val, nseg, out, long_list.p, nlong.p);
    IB_LAUNCHED();
    seg_sum_long_kernel<<<(unsigned)(2 * sm_count()), 256, 0, s>>>(ptr, perm, val, out, long_list.p, nlong.p);
    IB_LAUNCHED();
}


}   // namespace ib
