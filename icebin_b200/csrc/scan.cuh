// scan.cuh -- hand-written device-wide exclusive scan over a virtual input sequence.
//
//   transform_scan(n, load, op, ident, tmp..)        : phase 1+2 -- per-tile reduce, then scan of the tile sums
//   scan_apply(n, load, op, ident, apply)            : phase 3   -- per-tile rescan, apply(i, exclusive_prefix, x_i)
//
// `load(i)` produces element i on the fly (it is re-evaluated in phase 3, so it must be cheap and pure);
// several apply functors may be run against the same phase-1/2 result (the matrix build assigns dense ids with
// one and emits triplets with the next, with the grid-wide dependency carried by the kernel boundary).
// HBM traffic: whatever `load` reads, twice, plus 2*sizeof(T) per 1024-element tile.
#pragma once
#include "common.cuh"

namespace ib {

constexpr int SCAN_BLOCK = 256;
constexpr int SCAN_ITEMS = 4;
constexpr int SCAN_TILE = SCAN_BLOCK * SCAN_ITEMS;

template <class T>
__device__ __forceinline__ T shfl_up_any(T v, int delta)
{
    static_assert(sizeof(T) % 4 == 0, "T must be a multiple of 4 bytes");
    constexpr int W = sizeof(T) / 4;
    union { T t; uint32_t w[W]; } a, b;
    a.t = v;
#pragma unroll
    for (int k = 0; k < W; ++k) b.w[k] = __shfl_up_sync(0xffffffffu, a.w[k], delta);
    return b.t;
}

// Block-wide exclusive scan of one value per thread; returns the exclusive prefix, `total` = block total.
template <class T, class Op>
__device__ __forceinline__ T block_exclusive_scan(T v, Op op, T ident, T &total, T *smem /* [SCAN_BLOCK/32 + 1] */)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        T o = shfl_up_any(inc, d);
        if (lane >= d) inc = op(o, inc);
    }
    if (lane == 31) smem[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        constexpr int NW = SCAN_BLOCK / 32;
        T w = lane < NW ? smem[lane] : ident;
#pragma unroll
        for (int d = 1; d < NW; d <<= 1) {
            T o = shfl_up_any(w, d);
            if (lane >= d) w = op(o, w);
        }
        if (lane < NW) smem[lane] = w;     // inclusive over warps
    }
    __syncthreads();
    T excl = shfl_up_any(inc, 1);
    if (lane == 0) excl = ident;
    if (warp > 0) excl = op(smem[warp - 1], excl);
    total = smem[SCAN_BLOCK / 32 - 1];
    __syncthreads();
    return excl;
}

template <class T, class Op, class Load>
__global__ void __launch_bounds__(SCAN_BLOCK) scan_tile_reduce_kernel(int64_t n, Load load, Op op, T ident, T *tile_sums)
{
    __shared__ T smem[SCAN_BLOCK / 32 + 1];
    int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    T acc = ident;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        int64_t i = base + k;
        if (i < n) acc = op(acc, load(i));
    }
    T total;
    block_exclusive_scan(acc, op, ident, total, smem);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// Single block: exclusive scan of tile_sums in place; grand total to *total_out.
template <class T, class Op>
__global__ void __launch_bounds__(SCAN_BLOCK) scan_tile_sums_kernel(int64_t ntiles, Op op, T ident, T *tile_sums, T *total_out)
{
    __shared__ T smem[SCAN_BLOCK / 32 + 1];
    T carry = ident;
    for (int64_t start = 0; start < ntiles; start += SCAN_TILE) {
        int64_t base = start + (int64_t)threadIdx.x * SCAN_ITEMS;
        T x[SCAN_ITEMS];
        T acc = ident;
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; ++k) {
            x[k] = (base + k < ntiles) ? tile_sums[base + k] : ident;
            acc = op(acc, x[k]);
        }
        T total;
        T excl = block_exclusive_scan(acc, op, ident, total, smem);
        T run = op(carry, excl);
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; ++k) {
            if (base + k < ntiles) tile_sums[base + k] = run;
            run = op(run, x[k]);
        }
        carry = op(carry, total);
    }
    if (threadIdx.x == 0) *total_out = carry;
}

template <class T, class Op, class Load, class Apply>
__global__ void __launch_bounds__(SCAN_BLOCK) scan_apply_kernel(int64_t n, Load load, Op op, T ident, const T *tile_prefix, Apply apply)
{
    __shared__ T smem[SCAN_BLOCK / 32 + 1];
    int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    T x[SCAN_ITEMS];
    T acc = ident;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        x[k] = (base + k < n) ? load(base + k) : ident;
        acc = op(acc, x[k]);
    }
    T total;
    T excl = block_exclusive_scan(acc, op, ident, total, smem);
    T run = op(tile_prefix[blockIdx.x], excl);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        if (base + k < n) apply(base + k, run, x[k]);
        run = op(run, x[k]);
    }
}

// Host drivers ------------------------------------------------------------------------------------
template <class T>
struct ScanState {
    DBuf<T> tile_prefix;   // [ntiles] exclusive prefix of each tile
    DBuf<T> total;         // [1]
    int64_t n = 0, ntiles = 0;
};

template <class T, class Op, class Load>
void transform_scan(ScanState<T> &st, int64_t n, Load load, Op op, T ident, cudaStream_t s)
{
    st.n = n;
    st.ntiles = cdiv(n, SCAN_TILE);
    if (st.ntiles == 0) st.ntiles = 1;
    st.tile_prefix.alloc(st.ntiles, s);
    st.total.alloc(1, s);
    scan_tile_reduce_kernel<T, Op, Load><<<(unsigned)st.ntiles, SCAN_BLOCK, 0, s>>>(n, load, op, ident, st.tile_prefix.p);
    IB_LAUNCHED();
    scan_tile_sums_kernel<T, Op><<<1, SCAN_BLOCK, 0, s>>>(st.ntiles, op, ident, st.tile_prefix.p, st.total.p);
    IB_LAUNCHED();
}

template <class T, class Op, class Load, class Apply>
void scan_apply(ScanState<T> const &st, Load load, Op op, T ident, Apply apply, cudaStream_t s)
{
    scan_apply_kernel<T, Op, Load, Apply><<<(unsigned)st.ntiles, SCAN_BLOCK, 0, s>>>(st.n, load, op, ident, st.tile_prefix.p, apply);
    IB_LAUNCHED();
}

struct AddU32x4 {
    __host__ __device__ uint4 operator()(uint4 a, uint4 b) const { return make_uint4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
};
struct AddU32 {
    __host__ __device__ uint32_t operator()(uint32_t a, uint32_t b) const { return a + b; }
};

}   // namespace ib
