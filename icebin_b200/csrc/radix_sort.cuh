// radix_sort.cuh -- hand-written stable LSD radix sort of (uint64 key, payload) records on the device.
//
// Replaces spsparse::TupleList accumulation + Eigen::setFromTriplets' internal ordering
// (RegridMatrices_Dynamic.cpp:75-81 etc.): triplets are keyed row_dense << cbits | col_dense and sorted so that a
// reduce-by-key yields CSR.  8-bit digits; only ceil(keybits/8) passes run.  Each pass is
//   (1) per-tile digit histogram, (2) exclusive scan of the [digit][tile] table, (3) stable scatter with
//   warp-level match_any ranking.
// The sort is STABLE, so duplicate (row,col) triplets stay in emission order and are later summed in the same
// order Eigen sums them (input order).  HBM traffic per pass: 2 key reads + 1 payload read + 1 record write.
#pragma once
#include "common.cuh"
#include "scan.cuh"

namespace ib {

constexpr int RS_BLOCK = 256;
constexpr int RS_WARPS = RS_BLOCK / 32;
constexpr int RS_ITEMS = 8;
constexpr int RS_TILE = RS_BLOCK * RS_ITEMS;
constexpr int RS_RADIX = 256;

template <class K>
__global__ void __launch_bounds__(RS_BLOCK) rs_hist_kernel(const K *__restrict__ keys, int64_t n, int shift,
                                                           uint32_t *__restrict__ tile_hist, int ntiles)
{
    __shared__ uint32_t hist[RS_RADIX];
    hist[threadIdx.x] = 0;
    __syncthreads();
    int64_t base = (int64_t)blockIdx.x * RS_TILE;
#pragma unroll
    for (int k = 0; k < RS_ITEMS; ++k) {
        int64_t i = base + k * RS_BLOCK + threadIdx.x;
        if (i < n) atomicAdd(&hist[(uint32_t)(keys[i] >> shift) & (RS_RADIX - 1)], 1u);
    }
    __syncthreads();
    tile_hist[(int64_t)threadIdx.x * ntiles + blockIdx.x] = hist[threadIdx.x];
}

template <class K, class V>
__global__ void __launch_bounds__(RS_BLOCK) rs_scatter_kernel(const K *__restrict__ keys, const V *__restrict__ vals,
                                                              K *__restrict__ keys_out, V *__restrict__ vals_out,
                                                              int64_t n, int shift,
                                                              const uint32_t *__restrict__ tile_offs, int ntiles)
{
    __shared__ uint32_t wcount[RS_WARPS][RS_RADIX];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int d = lane; d < RS_RADIX; d += 32) wcount[warp][d] = 0;
    __syncwarp();

    const int64_t wbase = (int64_t)blockIdx.x * RS_TILE + (int64_t)warp * (RS_ITEMS * 32);
    K key[RS_ITEMS];
    uint32_t rank[RS_ITEMS];
    const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        int64_t i = wbase + r * 32 + lane;
        bool valid = i < n;
        key[r] = valid ? keys[i] : K(0);
        uint32_t d = valid ? ((uint32_t)(key[r] >> shift) & (RS_RADIX - 1)) : 0xffffffffu;
        uint32_t peers = __match_any_sync(0xffffffffu, d);
        int leader = __ffs(peers) - 1;
        uint32_t base = 0;
        if (lane == leader && valid) {
            base = wcount[warp][d];
            wcount[warp][d] = base + __popc(peers);
        }
        base = __shfl_sync(0xffffffffu, base, leader);
        rank[r] = base + __popc(peers & lt);
        __syncwarp();
    }
    __syncthreads();
    {   // digit `threadIdx.x`: turn per-warp counts into global bases
        int d = threadIdx.x;
        uint32_t run = tile_offs[(int64_t)d * ntiles + blockIdx.x];
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) {
            uint32_t c = wcount[w][d];
            wcount[w][d] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        int64_t i = wbase + r * 32 + lane;
        if (i < n) {
            uint32_t d = (uint32_t)(key[r] >> shift) & (RS_RADIX - 1);
            uint32_t pos = wcount[warp][d] + rank[r];
            keys_out[pos] = key[r];
            vals_out[pos] = vals[i];
        }
    }
}

// Exclusive scan of a small [digit][tile] table in ONE block (replaces three launches of the general scan when the table
// has at most RS_SMALL entries: sorts of up to ~0.5 M records)
constexpr int RS_SMALL_BLOCK = 1024;
constexpr int RS_SMALL = RS_SMALL_BLOCK * 64;
static __global__ void __launch_bounds__(RS_SMALL_BLOCK) rs_scan_small_kernel(uint32_t *hist, int n)
{
    __shared__ uint32_t wsum[RS_SMALL_BLOCK / 32];
    const int per = (n + RS_SMALL_BLOCK - 1) / RS_SMALL_BLOCK;
    const int b = threadIdx.x * per, e = min(n, b + per);
    uint32_t s = 0;
    for (int i = b; i < e; ++i) s += hist[i];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = s;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += o;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = wsum[lane];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(0xffffffffu, w, d);
            if (lane >= d) w += o;
        }
        wsum[lane] = w;
    }
    __syncthreads();
    uint32_t run = inc - s + (warp ? wsum[warp - 1] : 0u);
    for (int i = b; i < e; ++i) {
        const uint32_t c = hist[i];
        hist[i] = run;
        run += c;
    }
}

struct U32Load {
    const uint32_t *p;
    __device__ uint32_t operator()(int64_t i) const { return p[i]; }
};
struct U32Store {
    uint32_t *p;
    __device__ void operator()(int64_t i, uint32_t prefix, uint32_t) const { p[i] = prefix; }
};

// Sorts n records by the low `keybits` bits of the key.  (k0,v0) hold the input; (k1,v1) are scratch of the
// same size.  Returns 0 if the sorted data ends in (k0,v0), 1 if in (k1,v1).
template <class K, class V>
int radix_sort_pairs(K *k0, V *v0, K *k1, V *v1, int64_t n, int keybits, cudaStream_t s, int first_bit = 0)
{
    // first_bit > 0: the input is already ordered by the low `first_bit` bits; only the bits above them are sorted on
    // (a stable sort by the high bits of a stream sorted by the low bits is sorted by the whole key)
    keybits -= first_bit;
    if (n <= 1 || keybits <= 0) return 0;
    if (n >= (int64_t(1) << 32)) fail(IB_ERR_ARG, "radix_sort_pairs: n too large");
    int passes = (keybits + 7) / 8;
    int ntiles = (int)cdiv(n, RS_TILE);
    DBuf<uint32_t> hist((size_t)RS_RADIX * ntiles, s);
    ScanState<uint32_t> st;
    int cur = 0;
    for (int p = 0; p < passes; ++p) {
        K *ki = cur ? k1 : k0, *ko = cur ? k0 : k1;
        V *vi = cur ? v1 : v0, *vo = cur ? v0 : v1;
        rs_hist_kernel<K><<<ntiles, RS_BLOCK, 0, s>>>(ki, n, first_bit + p * 8, hist.p, ntiles);
        IB_LAUNCHED();
        if ((int64_t)RS_RADIX * ntiles <= RS_SMALL) {
            rs_scan_small_kernel<<<1, RS_SMALL_BLOCK, 0, s>>>(hist.p, RS_RADIX * ntiles);
            IB_LAUNCHED();
        } else {
            transform_scan<uint32_t>(st, (int64_t)RS_RADIX * ntiles, U32Load{hist.p}, AddU32{}, 0u, s);
            scan_apply<uint32_t>(st, U32Load{hist.p}, AddU32{}, 0u, U32Store{hist.p}, s);
        }
        rs_scatter_kernel<K, V><<<ntiles, RS_BLOCK, 0, s>>>(ki, vi, ko, vo, n, first_bit + p * 8, hist.p, ntiles);
        IB_LAUNCHED();
        cur ^= 1;
    }
    return cur;
}

}   // namespace ib
