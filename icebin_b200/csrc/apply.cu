// apply.cu -- batched fp64 CSR SpMM  Y[nrow x nvar] = M * X[ncol x nvar]  with fields stored cell-major
// (the nvar values of one cell are contiguous), so every nonzero drives one contiguous row read of X.
//
// Replaces ibmisc linear::Weighted_Eigen::apply()/apply_M() (call sites merge_topo.cpp:65, modele/icebin22m.cpp:153)
// and the explicit Eigen products of the coupling step (IceCoupler.cpp:237-238, 445-446).
//
// Work decomposition ("items", built once per matrix on the device):
//   * consecutive SHORT rows (<= APPLY_CHUNK nonzeros) are packed into groups of < 2*APPLY_CHUNK nonzeros and
//     <= 32 rows, so that matrices with 2 nonzeros per row (IvE, IvA) move as much data per unit as long rows do;
//   * LONG rows are cut into chunks of APPLY_CHUNK nonzeros whose partial sums are combined, in chunk order, by a
//     small second kernel (a fused "last block adds" variant was measured 15-25 % slower: its fence + atomic round
//     trip stalls the consumers of every chunk while the ring drains).
// Main kernel (spmm_tma_kernel): one producer warp streams the X rows named by an item's column indices into a
// shared-memory ring with 1-D bulk async copies (cp.async.bulk, SASS UBLKCP) completing on mbarriers; eight
// consumer warps read each landed row once (one double2 per thread) and keep running sums in registers.  Bytes in
// flight per SM are set by the ring, not by registers.  A plain-load kernel (spmm_ldg_kernel) covers small or odd
// nvar and unaligned buffers.
//
// Compiled with -fmad=false: a row is accumulated as acc = acc + v*x over ascending columns, exactly the order and
// rounding of Eigen's column-major sparse * dense on the CPU, so rows of <= APPLY_CHUNK nonzeros agree with the
// oracle bit for bit; split rows differ only in association.
//
// Roofline: HBM.  Algorithmic bytes = 12*nnz + 4*(nrow+1) + 8*nvar*(ncol_touched + nrow)   (SURVEY 8d).
#include <algorithm>
#include <mutex>

#include "gc_plan.cuh"
#include "radix_sort.cuh"
#include "scan.cuh"
#include "tma.cuh"
#include "types.cuh"
#include "util.cuh"

namespace ib {

constexpr int APPLY_CHUNK = 32;
constexpr int GROUP_ROWS = 32;

// ---- plan -----------------------------------------------------------------------------------------------
// item = (first_row, nnz_begin, nnz_end, w):  w >= 0: chunk of a long row, w = partial slot
//                                             w <  0: group of -w consecutive short rows
struct PlanLoad {
    const int32_t *rowptr;
    __device__ __forceinline__ int len(int64_t r) const { return rowptr[r + 1] - rowptr[r]; }
    __device__ __forceinline__ bool same_group(int64_t a, int64_t b) const   // a, b short rows, b = a + 1
    {
        return (rowptr[a] / APPLY_CHUNK == rowptr[b] / APPLY_CHUNK) && (a / GROUP_ROWS == b / GROUP_ROWS);
    }
    __device__ uint4 operator()(int64_t r) const
    {
        const int n = len(r);
        if (n > APPLY_CHUNK) {
            const uint32_t nch = (uint32_t)((n + APPLY_CHUNK - 1) / APPLY_CHUNK);
            return make_uint4(nch, nch, 1u, (uint32_t)n);
        }
        const bool head = (r == 0) || len(r - 1) > APPLY_CHUNK || !same_group(r - 1, r);
        return make_uint4(head ? 1u : 0u, 0u, 0u, (uint32_t)n);        // w: row length (combined with max)
    }
};
struct PlanApply {
    PlanLoad L;
    int64_t nrow;
    int4 *items;
    int4 *split_rows;      // (row, first partial slot, nchunks, 0)
    int32_t *slot_srow;    // partial slot -> split-row index
    __device__ void operator()(int64_t r, uint4 prefix, uint4 e) const
    {
        const int b = L.rowptr[r], end = L.rowptr[r + 1];
        if (e.z) {
            for (uint32_t q = 0; q < e.x; ++q) {
                const int lo = b + (int)q * APPLY_CHUNK;
                items[prefix.x + q] = make_int4((int)r, lo, min(end, lo + APPLY_CHUNK), (int)(prefix.y + q));
                slot_srow[prefix.y + q] = (int32_t)prefix.z;
            }
            split_rows[prefix.z] = make_int4((int)r, (int)prefix.y, (int)e.x, 0);
        } else if (e.x) {
            int64_t r2 = r + 1;
            while (r2 < nrow && L.len(r2) <= APPLY_CHUNK && L.same_group(r2 - 1, r2)) ++r2;
            items[prefix.x] = make_int4((int)r, b, L.rowptr[r2], -(int)(r2 - r));
        }
    }
};

static void make_gc_plan(Weighted &W, cudaStream_t s);
// items, partial slots and split rows add up; the longest row is a maximum -- one scan for all four
struct PlanOp {
    __host__ __device__ uint4 operator()(uint4 a, uint4 b) const { return make_uint4(a.x + b.x, a.y + b.y, a.z + b.z, a.w > b.w ? a.w : b.w); }
};

static void make_plan(Weighted &W, cudaStream_t s)
{
    ApplyPlan &P = W.plan;
    if (P.ready) return;
    if (W.nrow > 0) {
        ScanState<uint4> st;
        const uint4 z = make_uint4(0, 0, 0, 0);
        PlanLoad L{W.rowptr.p};
        transform_scan<uint4>(st, W.nrow, L, PlanOp{}, z, s);
        uint4 tot;
        IB_CUDA(cudaMemcpyAsync(&tot, st.total.p, sizeof tot, cudaMemcpyDeviceToHost, s));
        IB_CUDA(cudaStreamSynchronize(s));
        P.nitems = tot.x; P.npartials = tot.y; P.nsplit_rows = tot.z; P.max_row_len = (int)tot.w;
        P.items.alloc((size_t)P.nitems, s);
        P.split_rows.alloc((size_t)P.nsplit_rows, s);
        P.slot_srow.alloc((size_t)P.npartials, s);
        scan_apply<uint4>(st, L, PlanOp{}, z, PlanApply{L, W.nrow, P.items.p, P.split_rows.p, P.slot_srow.p}, s);
        if (P.max_row_len <= GC_S && P.nsplit_rows == 0 && g_tune.use_gc) make_gc_plan(W, s);
    }
    P.ready = true;
}

// ---- distinct-column groups (gc_plan.cuh): the greedy grouping of a 32-row block is sequential, but its inputs (33 row
// pointers, <= 256 columns and values) are staged into shared memory by the whole warp with coalesced loads first; then
// lane 0 runs gc_block on the staged copy.  Twice: count -> scan -> emit.  (One thread per block straight from global
// memory spent 1.4 ms on 2149 blocks at 5 km: a chain of dependent loads per row.)
constexpr int GCP_WARPS = 4;
template <bool EMIT>
__global__ void __launch_bounds__(GCP_WARPS * 32) gc_plan_kernel(const int32_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                                                                 const double *__restrict__ val, int64_t nrow, int64_t nblk,
                                                                 uint2 *__restrict__ counts, const uint4 *__restrict__ prefix,
                                                                 GcItem *items, int32_t *ucol, GcRec *recs)
{
    __shared__ int32_t s_rp[GCP_WARPS][GC_BLOCK + 1];
    __shared__ int32_t s_col[GCP_WARPS][GC_BLOCK * GC_S];
    __shared__ double s_val[GCP_WARPS][GC_BLOCK * GC_S];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t b = (int64_t)blockIdx.x * GCP_WARPS + w;
    if (b >= nblk) return;
    const int64_t r0 = b * GC_BLOCK, r1 = min(nrow, r0 + GC_BLOCK);
    for (int k = lane; k <= (int)(r1 - r0); k += 32) s_rp[w][k] = __ldg(rowptr + r0 + k);
    __syncwarp();
    const int32_t j0 = s_rp[w][0], j1 = s_rp[w][r1 - r0];
    for (int32_t j = j0 + lane; j < j1; j += 32) {           // j1 - j0 <= 256: every row has <= GC_S nonzeros
        s_col[w][j - j0] = __ldg(col + j);
        if (EMIT) s_val[w][j - j0] = __ldg(val + j);
    }
    __syncwarp();
    if (lane == 0) {
        // gc_block indexes rowptr[r] with r in [r0, r1] and col/val[j] with j in [j0, j1): hand it rebased views
        const int32_t *rp = s_rp[w] - r0, *cc = s_col[w] - j0;
        const double *vv = s_val[w] - j0;
        if (EMIT) {
            const uint4 p = prefix[b];
            gc_block<true>(rp, cc, vv, nrow, b, p.x, p.y, items, ucol, recs);
        } else {
            const GcCount c = gc_block<false>(rp, cc, nullptr, nrow, b, 0u, 0u, nullptr, nullptr, nullptr);
            counts[b] = make_uint2(c.items, c.recs);
        }
    }
}
struct GcCountLoad {
    const uint2 *counts;
    __device__ uint4 operator()(int64_t b) const { const uint2 c = counts[b]; return make_uint4(c.x, c.y, 0u, 0u); }
};
struct GcPrefixStore {
    uint4 *prefix;
    __device__ void operator()(int64_t b, uint4 p, uint4) const { prefix[b] = p; }
};
static_assert(sizeof(GcItem) == sizeof(int4) && sizeof(GcRec) == sizeof(int4), "GcItem / GcRec must alias int4");

static void make_gc_plan(Weighted &W, cudaStream_t s)
{
    ApplyPlan &P = W.plan;
    const int64_t nblk = cdiv(W.nrow, GC_BLOCK);
    DBuf<uint2> counts((size_t)nblk, s);
    DBuf<uint4> prefix((size_t)nblk, s);
    const unsigned grid = (unsigned)cdiv(nblk, GCP_WARPS);
    gc_plan_kernel<false><<<grid, GCP_WARPS * 32, 0, s>>>(W.rowptr.p, W.col.p, W.val.p, W.nrow, nblk, counts.p, nullptr, nullptr, nullptr, nullptr);
    IB_LAUNCHED();
    ScanState<uint4> st;
    const uint4 z = make_uint4(0, 0, 0, 0);
    GcCountLoad L{counts.p};
    transform_scan<uint4>(st, nblk, L, AddU32x4{}, z, s);
    scan_apply<uint4>(st, L, AddU32x4{}, z, GcPrefixStore{prefix.p}, s);
    uint4 tot;
    IB_CUDA(cudaMemcpyAsync(&tot, st.total.p, sizeof tot, cudaMemcpyDeviceToHost, s));
    IB_CUDA(cudaStreamSynchronize(s));
    if ((int64_t)tot.y != W.nnz) return;          // empty rows (dummy records): left to the other kernels
    P.gc_nitems = tot.x;
    P.gc_items.alloc((size_t)tot.x, s);
    P.gc_ucol.alloc((size_t)tot.x * GC_S, s);
    P.gc_recs.alloc((size_t)tot.y + 2, s);         // the kernel's record prefetch runs two entries ahead
    P.gc_recs.zero();
    P.gc_counter.alloc(3, s);
    P.gc_counter.zero();
    gc_plan_kernel<true><<<grid, GCP_WARPS * 32, 0, s>>>(W.rowptr.p, W.col.p, W.val.p, W.nrow, nblk, nullptr, prefix.p,
                                                         reinterpret_cast<GcItem *>(P.gc_items.p), P.gc_ucol.p, reinterpret_cast<GcRec *>(P.gc_recs.p));
    IB_LAUNCHED();
    IB_CUDA(cudaStreamSynchronize(s));           // counts / prefix die here
    P.gc = true;
}

__global__ void expand_rows_kernel(const int32_t *rowptr, int64_t nrow, int32_t *row)
{
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrow) return;
    for (int32_t j = rowptr[r]; j < rowptr[r + 1]; ++j) row[j] = (int32_t)r;
}
__global__ void u32_to_i32_copy_kernel(const uint32_t *a, int64_t n, int32_t *b)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) b[i] = (int32_t)a[i];
}

// ---- TMA (cp.async.bulk) pipeline -----------------------------------------------------------------------
constexpr int TMA_ROWS = 4;          // X rows per pipeline stage (amortises the barrier round trip)
constexpr int TMA_CONSUMERS = 256;   // consumer threads (8 warps); +1 producer warp
constexpr int TMA_MAX_TW = 2 * TMA_CONSUMERS;   // fields per tile

struct SpmmArgs {
    const int4 *items;
    int64_t nunits;
    int ntile, tw, nstage;
    const int32_t *rowptr, *col;
    const double *val, *X;
    int64_t ldx;
    int nvar;
    double *Y;
    int64_t ldy;
    double *partials;
    int64_t ldp;
    int row_lo, row_hi;           // only items whose first row lies in [row_lo, row_hi) are processed (row-block sharding)
};

// Shared-memory stage = [TMA_ROWS rows of X][header: TMA_ROWS values]; the producer also delivers the matrix values of
// the stage through the header, so consumers touch global memory only to read the (prefetched) item record and to
// write results.
constexpr uint32_t TMA_HDR = 2 * TMA_ROWS * 8;

__global__ void __launch_bounds__(TMA_CONSUMERS + 32) spmm_tma_kernel(const SpmmArgs a)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw);
    uint64_t *empty = full + a.nstage;
    unsigned char *data = smem_raw + 256;           // 2*nstage*8 <= 248
    const uint32_t rowbytes = (uint32_t)a.tw * 8u;
    const uint32_t stagebytes = rowbytes * TMA_ROWS + TMA_HDR;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < a.nstage; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, TMA_CONSUMERS / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    int stage = 0;
    uint32_t phase = 0;
    const int64_t ustep = (int64_t)gridDim.x;
    int64_t u = (int64_t)blockIdx.x;
    const int64_t uend = a.nunits;
    int4 it_next = u < uend ? a.items[u / a.ntile] : make_int4(0, 0, 0, 0);
    if (warp == TMA_CONSUMERS / 32) {
        // ===== producer warp: lanes 0..nr-1 each issue one row copy of the stage and post its matrix value
        for (; u < uend; u += ustep) {
            const int64_t item = u / a.ntile;
            const int tile = (int)(u - item * a.ntile);
            const int4 it = it_next;
            if (u + ustep < uend) it_next = a.items[(u + ustep) / a.ntile];
            if (it.x < a.row_lo || it.x >= a.row_hi) continue;
            const int f0 = tile * a.tw;
            const uint32_t tb = (uint32_t)min(a.tw, a.nvar - f0) * 8u;
            for (int j = it.y; j < it.z; j += TMA_ROWS) {
                const int nr = min(TMA_ROWS, it.z - j);
                int c = 0;
                double v = 0.0;
                if (lane < nr) {
                    c = __ldg(a.col + j + lane);
                    v = __ldg(a.val + j + lane);
                }
                if (lane == 0) mbar_wait(empty + stage, phase ^ 1u);
                __syncwarp();
                unsigned char *sp = data + (size_t)stage * stagebytes;
                if (lane < TMA_ROWS) {
                    double *hdr = reinterpret_cast<double *>(sp + (size_t)TMA_ROWS * rowbytes);
                    hdr[lane] = v;
                }
                __syncwarp();
                if (lane == 0) mbar_expect_tx(full + stage, tb * (uint32_t)nr);      // arrive (release) after the header stores
                __syncwarp();
                if (lane < nr) bulk_g2s(sp + (size_t)lane * rowbytes, a.X + (int64_t)c * a.ldx + f0, tb, full + stage);
                if (++stage == a.nstage) { stage = 0; phase ^= 1u; }
            }
        }
    } else {
        // ===== consumer warps: thread t owns fields f0 + 2t, f0 + 2t + 1
        const int t = threadIdx.x;
        for (; u < uend; u += ustep) {
            const int64_t item = u / a.ntile;
            const int tile = (int)(u - item * a.ntile);
            const int4 it = it_next;
            if (u + ustep < uend) it_next = a.items[(u + ustep) / a.ntile];
            if (it.x < a.row_lo || it.x >= a.row_hi) continue;
            const int f0 = tile * a.tw;
            const bool active = 2 * t < min(a.tw, a.nvar - f0);
            const int nrows = it.w < 0 ? -it.w : 0;
            // row ends of a group live one per lane (nrows <= 32)
            int my_end = (lane < nrows) ? __ldg(a.rowptr + it.x + lane + 1) : 0;
            int cur = 0;
            int cur_end = nrows ? __shfl_sync(0xffffffffu, my_end, 0) : it.z;
            double a0 = 0.0, a1 = 0.0;
            int j = it.y;
            // leading empty rows of a group
            while (cur < nrows && cur_end == j) {
                if (active) reinterpret_cast<double2 *>(a.Y + (int64_t)(it.x + cur) * a.ldy + f0)[t] = make_double2(0.0, 0.0);
                ++cur;
                cur_end = __shfl_sync(0xffffffffu, my_end, cur & 31);
            }
            for (int jb = it.y; jb < it.z; jb += TMA_ROWS) {
                const int nr = min(TMA_ROWS, it.z - jb);
                mbar_wait(full + stage, phase);
                const unsigned char *sb = data + (size_t)stage * stagebytes;
                const double *hdr = reinterpret_cast<const double *>(sb + (size_t)TMA_ROWS * rowbytes);
                const unsigned char *sp = sb + (size_t)t * 16;
#pragma unroll
                for (int r = 0; r < TMA_ROWS; ++r) {
                    if (r < nr) {
                        if (active) {
                            const double2 x = *reinterpret_cast<const double2 *>(sp + (size_t)r * rowbytes);
                            const double v = hdr[r];
                            a0 = a0 + v * x.x;
                            a1 = a1 + v * x.y;
                        }
                        ++j;
                        while (cur < nrows && cur_end == j) {      // row(s) of a group complete
                            if (active) reinterpret_cast<double2 *>(a.Y + (int64_t)(it.x + cur) * a.ldy + f0)[t] = make_double2(a0, a1);
                            a0 = 0.0; a1 = 0.0;
                            ++cur;
                            cur_end = __shfl_sync(0xffffffffu, my_end, cur & 31);
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(empty + stage);
                if (++stage == a.nstage) { stage = 0; phase ^= 1u; }
            }
            // chunk of a long row: partial sum, combined in chunk order by the fix-up
            if (it.w >= 0 && active) reinterpret_cast<double2 *>(a.partials + (int64_t)it.w * a.ldp + f0)[t] = make_double2(a0, a1);
        }
    }
}

// ---- X-stationary kernel: tall matrices with few columns (IvE, IvA, smoothed IvE) ------------------------------
// When the input space is small (E or A cells: ~10^3) and the output space large (ice cells: 10^5..10^7), a field slice
// of ALL of X fits in shared memory: Xs[ncol][F] with F = 8..64 fields.  A CTA keeps one slice resident and streams
// matrix rows through it: thread (row, p) walks its row's nonzeros, reads X from shared memory and writes one double2
// of Y -- X costs no L2 traffic per nonzero (the row-streaming kernel re-fetches a 4 KB X row per nonzero, which is what
// made IvE 0.42 and the 16-nnz/row smoothed IvE L2-bound).  The price is that the matrix is re-read once per slice.
// Rows are never split: every row is summed over ascending columns with mul-then-add, i.e. bit-identical to the oracle.
struct XsArgs {
    const int32_t *rowptr, *col;
    const double *val, *X;
    int64_t ldx;
    double *Y;
    int64_t ldy;
    int64_t nrow;
    int ncol, nvar, F, nslice;
    int64_t nchunk, nunits;
    int chunk_rows;
};

constexpr int XS_THREADS = 1024;
constexpr int XS_K = 4;          // rows in flight per thread (independent load chains)

__global__ void __launch_bounds__(XS_THREADS) spmm_xs_kernel(const XsArgs a)
{
    extern __shared__ __align__(16) unsigned char xs_raw[];
    double2 *Xs = reinterpret_cast<double2 *>(xs_raw);
    const int TPR = a.F >> 1;                       // threads (double2) per row
    const int rpp = blockDim.x / TPR;               // rows per pass
    const int p = threadIdx.x % TPR, rl = threadIdx.x / TPR;
    // contiguous share of the (slice-major) unit list: a CTA changes slice at most a couple of times
    const int64_t u0 = a.nunits * blockIdx.x / gridDim.x, u1 = a.nunits * (blockIdx.x + 1) / gridDim.x;
    int cur_slice = -1;
    for (int64_t u = u0; u < u1; ++u) {
        const int slice = (int)(u / a.nchunk);
        const int64_t chunk = u - (int64_t)slice * a.nchunk;
        const int f0 = slice * a.F;
        const int fw = min(a.F, a.nvar - f0) >> 1;  // live double2 per row in this slice
        if (slice != cur_slice) {
            __syncthreads();
            for (int idx = threadIdx.x; idx < a.ncol * TPR; idx += blockDim.x) {
                const int c = idx / TPR, q = idx - c * TPR;
                Xs[idx] = q < fw ? __ldg(reinterpret_cast<const double2 *>(a.X + (int64_t)c * a.ldx + f0) + q) : make_double2(0.0, 0.0);
            }
            __syncthreads();
            cur_slice = slice;
        }
        const int64_t r_end = min(a.nrow, (chunk + 1) * a.chunk_rows);
        for (int64_t rb = chunk * a.chunk_rows + rl; rb < r_end; rb += (int64_t)rpp * XS_K) {
            int jb[XS_K], je[XS_K];
#pragma unroll
            for (int k = 0; k < XS_K; ++k) {
                const int64_t r = rb + (int64_t)k * rpp;
                const bool ok = r < r_end;
                jb[k] = ok ? __ldg(a.rowptr + r) : 0;
                je[k] = ok ? __ldg(a.rowptr + r + 1) : 0;
            }
            // first two nonzeros of every row up front (IvE / IvA rows have 1..4): independent loads
            int c0[XS_K], c1[XS_K];
            double v0[XS_K], v1[XS_K];
#pragma unroll
            for (int k = 0; k < XS_K; ++k) {
                const bool h0 = jb[k] < je[k], h1 = jb[k] + 1 < je[k];
                c0[k] = h0 ? __ldg(a.col + jb[k]) : 0;
                v0[k] = h0 ? __ldg(a.val + jb[k]) : 0.0;
                c1[k] = h1 ? __ldg(a.col + jb[k] + 1) : 0;
                v1[k] = h1 ? __ldg(a.val + jb[k] + 1) : 0.0;
            }
#pragma unroll
            for (int k = 0; k < XS_K; ++k) {
                const int64_t r = rb + (int64_t)k * rpp;
                if (r >= r_end) continue;
                double a0 = 0.0, a1 = 0.0;
                if (jb[k] < je[k]) { const double2 x = Xs[c0[k] * TPR + p]; a0 = a0 + v0[k] * x.x; a1 = a1 + v0[k] * x.y; }
                if (jb[k] + 1 < je[k]) { const double2 x = Xs[c1[k] * TPR + p]; a0 = a0 + v1[k] * x.x; a1 = a1 + v1[k] * x.y; }
                int j = jb[k] + 2;
                for (; j + 4 <= je[k]; j += 4) {
                    int cc[4];
                    double vv[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) { cc[q] = __ldg(a.col + j + q); vv[q] = __ldg(a.val + j + q); }
#pragma unroll
                    for (int q = 0; q < 4; ++q) { const double2 x = Xs[cc[q] * TPR + p]; a0 = a0 + vv[q] * x.x; a1 = a1 + vv[q] * x.y; }
                }
                for (; j < je[k]; ++j) {
                    const int cq = __ldg(a.col + j);
                    const double vq = __ldg(a.val + j);
                    const double2 x = Xs[cq * TPR + p];
                    a0 = a0 + vq * x.x; a1 = a1 + vq * x.y;
                }
                if (p < fw) reinterpret_cast<double2 *>(a.Y + r * a.ldy + f0)[p] = make_double2(a0, a1);
            }
        }
    }
}

// ---- distinct-column group kernel: short-row matrices (IvE, IvA, EvA) --------------------------------------------
// Unit = (item of gc_plan.cuh, tile of 64*NQ fields), owned by ONE WARP; lane l owns the double2 at fields
// tile + 64q + 2l, q < NQ.  The X rows of the item's <= GC_S distinct columns are fetched ONCE, by 1-D bulk async copies
// (cp.async.bulk, one per row, issued by lanes 0..nd-1) into a warp-private shared-memory slot table, together with the
// item's records (value, slot, flags: 16 bytes per nonzero, prepared once per matrix); tables are double buffered: the
// next unit lands (completing on a warp-private mbarrier) while the current one is processed.  The inner loop is a flat
// walk over the records: broadcast-read a record, read the X row of its slot from shared memory, multiply-add, and on
// the last record of a row store the row.  Record reads run two records ahead and X reads one record ahead of the
// arithmetic.  Warps never synchronise with each other.
// History: (1) one double2 per lane, values/slots broadcast by shuffles: bound by instruction issue (58 warp
// instructions per nonzero and 64 fields); (2) NQ double2 per lane: 160 instructions per row and tile, every one
// waiting ~6 cycles on the previous (3 shuffles + convergence checks per nonzero); (3) this one.
// X traffic drops from one row per nonzero to one per distinct column (IvE 0.15x, IvA 0.06x at 5 km), which is what
// bounded the row-streaming kernels: they ran at the L2 fabric rate (reads + writes ~ 8.4 TB/s), not at HBM's.
// Rows are never split and are summed over ascending columns with mul-then-add: bit-identical to the oracle.
struct GcArgs {
    const int4 *items;
    const int32_t *ucol;
    const int4 *recs;
    int64_t nitems;
    int ntile;
    const double *X;
    int64_t ldx;
    int nvar;
    double *Y;
    int64_t ldy;
    int row_lo, row_hi;
    unsigned long long *counter;     // [0] units handed out so far (never reset), [1] its value when this launch began, [2] warps done
};
constexpr int GC_WARPS = 2;      // warps per CTA (independent; a CTA is only a container for shared memory)
constexpr int GC_RECBUF = (GC_MAXREC + 2) * 16;     // record table of one buffer (+2: the prefetch overruns)

template <int NQ>
constexpr size_t gc_smem_bytes() { return (size_t)GC_WARPS * (2 * (GC_S * 64 * NQ * 8 + GC_RECBUF) + 16); }

template <int NQ>
__global__ void __launch_bounds__(GC_WARPS * 32) spmm_gc_kernel(const GcArgs a)
{
    constexpr int TW = 64 * NQ;                        // fields per tile
    constexpr uint32_t ROWB = TW * 8;                  // bytes per slot
    constexpr uint32_t BUFB = GC_S * ROWB + GC_RECBUF; // one buffer: slot table + record table
    extern __shared__ __align__(128) unsigned char gc_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char *buf = gc_raw + (size_t)warp * (2 * BUFB);
    uint64_t *bar = reinterpret_cast<uint64_t *>(gc_raw + (size_t)GC_WARPS * (2 * BUFB)) + 2 * warp;
    const unsigned FULL = 0xffffffffu;
    if (lane == 0) {
        mbar_init(bar, 1);
        mbar_init(bar + 1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    const int64_t nunits = a.nitems * a.ntile;

    // Units are handed out dynamically (items differ in size: 1..32 rows): unit u = item * ntile + tile.  The counter is
    // read three units ahead of the arithmetic, so its latency is never waited for.
    // the first three units of a warp are static; the counter hands out the rest.  It is never reset: the last warp
    // to leave records where the next launch starts (so a captured graph can be replayed)
    const int64_t nw = (int64_t)gridDim.x * GC_WARPS;
    const unsigned long long cbase = a.counter[1] - 3ull * (unsigned long long)nw;
    auto take_unit = [&]() -> unsigned long long {          // lane 0 holds the answer; broadcast where it is used
        unsigned long long u = 0;
        if (lane == 0) u = atomicAdd(a.counter, 1ull) - cbase;
        return u;
    };
    auto bcast = [&](unsigned long long u) -> int64_t { return (int64_t)__shfl_sync(FULL, u, 0); };
    // record of a unit: the item and (lanes 0..GC_S-1) its distinct columns.  Nothing here waits for a load: the
    // row-block test is applied where the record is used.
    auto load_item = [&](int64_t u, int4 &it, int &uc) {
        it = make_int4(0, 0, 0, 0);
        uc = -1;
        if (u < nunits) {
            const int64_t item = (uint32_t)u / (uint32_t)a.ntile;        // nunits < 2^31 (checked by the host)
            it = __ldg(a.items + item);
            if (lane < GC_S) uc = __ldg(a.ucol + item * GC_S + lane);
        }
    };
    auto in_block = [&](const int4 &it) { return it.x >= a.row_lo && it.x < a.row_hi; };
    // X rows and records of a unit -> buffer b.  The buffer was last read by this warp two units ago (program order +
    // __syncwarp); units outside the row block fetch nothing.
    auto issue = [&](int64_t u, const int4 &it, int uc, int b) {
        const bool ok = in_block(it);
        const int f0 = (int)((uint32_t)u % (uint32_t)a.ntile) * TW;
        const uint32_t tb = (uint32_t)min(TW, a.nvar - f0) * 8u;
        const uint32_t rb = ok ? (uint32_t)(it.z - it.y) * 16u : 0u;
        const bool mine = ok && uc >= 0;                     // lanes >= GC_S carry uc = -1
        const int nd = __popc(__ballot_sync(FULL, mine));
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic reads of the buffer before async writes
        __syncwarp();
        unsigned char *dst = buf + (size_t)b * BUFB;
        if (lane == 0) {
            if (rb + nd > 0) mbar_expect_tx(bar + b, tb * (uint32_t)nd + rb);
            else mbar_arrive(bar + b);
        }
        __syncwarp();
        if (mine) bulk_g2s(dst + (size_t)lane * ROWB, a.X + (int64_t)uc * a.ldx + f0, tb, bar + b);
        if (lane == GC_S && rb > 0) bulk_g2s(dst + GC_S * ROWB, a.recs + it.y, rb, bar + b);
    };

    int64_t u_c = (int64_t)blockIdx.x * GC_WARPS + warp, u_n = u_c + nw, u_n2 = u_n + nw;
    unsigned long long raw_n3 = take_unit();
    int4 itc, itn, itn2;
    int ucc, ucn, ucn2;
    load_item(u_c, itc, ucc);
    load_item(u_n, itn, ucn);
    int b = 0;
    uint32_t phase = 0;                                      // bit b: parity to wait for on bar[b]
    if (u_c < nunits) issue(u_c, itc, ucc, 0);
    while (u_c < nunits) {
        // prefetch: tables of the next unit, item record of the one after, unit number of the one after that
        if (u_n < nunits) issue(u_n, itn, ucn, b ^ 1);
        load_item(u_n2, itn2, ucn2);
        const int64_t u_n3 = bcast(raw_n3);
        raw_n3 = take_unit();
        // this unit's tables have landed?
        mbar_wait(bar + b, (phase >> b) & 1u);
        phase ^= 1u << b;
        const int n = in_block(itc) ? itc.z - itc.y : 0;
        const int f0 = (int)((uint32_t)u_c % (uint32_t)a.ntile) * TW;
        const int nq = min(NQ, (a.nvar - f0 + 63) >> 6);            // live 64-field groups of this tile
        const bool last_live = f0 + 64 * (nq - 1) + 2 * lane < a.nvar;   // ragged last group
        bool live[NQ];
#pragma unroll
        for (int q = 0; q < NQ; ++q) live[q] = q < nq - 1 || (q == nq - 1 && last_live);
        const double2 *xb = reinterpret_cast<const double2 *>(buf + (size_t)b * BUFB) + lane;
        const int4 *rec = reinterpret_cast<const int4 *>(buf + (size_t)b * BUFB + GC_S * ROWB);
        double2 *yp = reinterpret_cast<double2 *>(a.Y + (int64_t)itc.x * a.ldy + f0) + lane;
        const int64_t ystep = a.ldy / 2;
        double2 acc[NQ], xA[NQ], xB[NQ];
#pragma unroll
        for (int q = 0; q < NQ; ++q) acc[q] = make_double2(0.0, 0.0);
        auto fetch = [&](double2(&x)[NQ], int slot) {
            const double2 *xr = xb + (slot & (GC_S - 1)) * (TW / 2);
#pragma unroll
            for (int q = 0; q < NQ; ++q) x[q] = xr[q * 32];
        };
        // branch-free: the row sum so far goes out through its own registers when the record closes its row (the
        // accumulators are not written while a store still reads them), and the accumulators restart from zero
        auto step = [&](const int4 &r, const double2(&x)[NQ]) {
            const double v = __hiloint2double(r.y, r.x);
            const bool last = (r.w & (int)GC_LAST) != 0;
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                double2 t;
                t.x = acc[q].x + v * x[q].x;
                t.y = acc[q].y + v * x[q].y;
                if (last && live[q]) yp[q * 32] = t;
                acc[q].x = last ? 0.0 : t.x;
                acc[q].y = last ? 0.0 : t.y;
            }
            yp += last ? ystep : 0;
        };
        // flat walk over the records, two per trip: a record names the slot of its successor, so the successor's X row
        // is fetched (and the successor read) while the record itself is processed
        int4 rA = rec[0], rB;
        fetch(xA, rA.z);
        for (int k = 0; k < n; k += 2) {
            rB = rec[k + 1];
            fetch(xB, rA.z >> 8);
            step(rA, xA);
            if (k + 1 >= n) break;
            rA = rec[k + 2];
            fetch(xA, rB.z >> 8);
            step(rB, xB);
        }
        __syncwarp();
        itc = itn; itn = itn2;
        ucn = ucn2;
        u_c = u_n; u_n = u_n2; u_n2 = u_n3;
        b ^= 1;
    }
    if (lane == 0) {
        __threadfence();
        if (atomicAdd(a.counter + 2, 1ull) == (unsigned long long)nw - 1) {      // every other warp has taken its last unit
            a.counter[1] = atomicAdd(a.counter, 0ull);
            a.counter[2] = 0;
        }
    }
}

// ---- plain-load fallback (small / odd nvar, unaligned buffers) ---------------------------------------------
// A group of L lanes (power of two) owns one (item, field-tile) unit; lane fl owns VEC consecutive fields.
template <int VEC>
__global__ void __launch_bounds__(256) spmm_ldg_kernel(const int4 *__restrict__ items, int64_t nunits, int ntile, int L,
                                                       const int32_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                                                       const double *__restrict__ val, const double *__restrict__ X, int64_t ldx,
                                                       int nvar, double *__restrict__ Y, int64_t ldy, double *__restrict__ partials,
                                                       int64_t ldp, int row_lo, int row_hi)
{
    const int lane = threadIdx.x & 31;
    const int G = 32 / L;
    const int64_t gwarp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int fl = lane % L;
    for (int64_t u = gwarp * G + lane / L; u < nunits; u += nwarps * G) {
        const int64_t item = u / ntile;
        const int tile = (int)(u - item * ntile);
        const int f0 = (tile * L + fl) * VEC;
        if (f0 >= nvar) continue;
        const int4 it = items[item];
        if (it.x < row_lo || it.x >= row_hi) continue;
        const int nrows = it.w < 0 ? -it.w : 1;
        for (int rr = 0; rr < nrows; ++rr) {
            const int jb = it.w < 0 ? __ldg(rowptr + it.x + rr) : it.y;
            const int je = it.w < 0 ? __ldg(rowptr + it.x + rr + 1) : it.z;
            double acc[VEC];
#pragma unroll
            for (int k = 0; k < VEC; ++k) acc[k] = 0.0;
            int j = jb;
            for (; j + 4 <= je; j += 4) {
                int c[4];
                double v[4];
                double x[4][VEC];
#pragma unroll
                for (int q = 0; q < 4; ++q) { c[q] = __ldg(col + j + q); v[q] = __ldg(val + j + q); }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const double *xp = X + (int64_t)c[q] * ldx + f0;
                    if (VEC == 2) { double2 tt = __ldg(reinterpret_cast<const double2 *>(xp)); x[q][0] = tt.x; x[q][VEC - 1] = tt.y; }
                    else x[q][0] = __ldg(xp);
                }
#pragma unroll
                for (int q = 0; q < 4; ++q)
#pragma unroll
                    for (int k = 0; k < VEC; ++k) acc[k] = acc[k] + v[q] * x[q][k];
            }
            for (; j < je; ++j) {
                const int c = __ldg(col + j);
                const double v = __ldg(val + j);
                const double *xp = X + (int64_t)c * ldx + f0;
                if (VEC == 2) { double2 tt = __ldg(reinterpret_cast<const double2 *>(xp)); acc[0] = acc[0] + v * tt.x; acc[VEC - 1] = acc[VEC - 1] + v * tt.y; }
                else acc[0] = acc[0] + v * __ldg(xp);
            }
            double *out = (it.w < 0) ? (Y + (int64_t)(it.x + rr) * ldy + f0) : (partials + (int64_t)it.w * ldp + f0);
            if (VEC == 2) *reinterpret_cast<double2 *>(out) = make_double2(acc[0], acc[VEC - 1]);
            else out[0] = acc[0];
        }
    }
}

// Y[row] = sum of the row's chunk partials, added in chunk order (deterministic)
template <int VEC>
__global__ void __launch_bounds__(256) spmm_fixup_kernel(const int4 *__restrict__ split_rows, int64_t nsplit,
                                                         const double *__restrict__ partials, int64_t ldp, int nvar,
                                                         double *__restrict__ Y, int64_t ldy, int row_lo, int row_hi)
{
    const int nv = nvar / VEC;
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t r = t / nv;
    int f = (int)(t - r * nv) * VEC;
    asm volatile("griddepcontrol.wait;" ::: "memory");      // launched with programmatic stream serialization
    if (r >= nsplit) return;
    const int4 sr = split_rows[r];
    if (sr.x < row_lo || sr.x >= row_hi) return;
    const double *pp = partials + (int64_t)sr.y * ldp + f;
    double s[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) s[k] = 0.0;   // 0 + p0 == p0 exactly
    int q = 0;
    for (; q + 8 <= sr.z; q += 8) {
        double p[8][VEC];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (VEC == 2) { double2 tt = __ldcg(reinterpret_cast<const double2 *>(pp + (int64_t)(q + u) * ldp)); p[u][0] = tt.x; p[u][VEC - 1] = tt.y; }
            else p[u][0] = __ldcg(pp + (int64_t)(q + u) * ldp);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int k = 0; k < VEC; ++k) s[k] = s[k] + p[u][k];
    }
    for (; q < sr.z; ++q)
#pragma unroll
        for (int k = 0; k < VEC; ++k) s[k] = s[k] + __ldcg(pp + (int64_t)q * ldp + k);
    double *out = Y + (int64_t)sr.x * ldy + f;
    if (VEC == 2) *reinterpret_cast<double2 *>(out) = make_double2(s[0], s[VEC - 1]);
    else out[0] = s[0];
}

// ---- conservation correction (ibmisc Weighted_Eigen::apply_e; matrix_formats.rst:156-165) ---------------
// The whole correction of the input in ONE launch: a block owns 4 fields (one 32-byte sector per row) and walks ALL
// rows twice.  Pass 1: thread (rg, fi) adds up rows rg, rg + 128, ... of field fi for both weight vectors (nine
// independent loads at 1120 rows, all in flight at once); the 128 row groups are added in a fixed order through shared
// memory (deterministic): factor[f] = (wa . A)[f] / (wb . A)[f].  Pass 2: As[r][f] = A[r][f] * factor[f] (the rows are
// still in L1 / L2).  128 blocks at 512 fields: one per SM.
__global__ void __launch_bounds__(512) cons_scale_kernel(const double *__restrict__ wa, const double *__restrict__ wb,
                                                         const double *__restrict__ A, int64_t lda, int64_t nrow, int nvar,
                                                         double *__restrict__ As, int64_t ldas)
{
    __shared__ double sa_s[128][4], sb_s[128][4];
    __shared__ double pa_s[8][4], pb_s[8][4];
    __shared__ double fac_s[4];
    const int rg = threadIdx.x >> 2, fi = threadIdx.x & 3;
    const int f = blockIdx.x * 4 + fi;
    pdl_launch_dependents();             // the apply that follows may set itself up; it waits before it reads As
    double sa = 0, sb = 0;
    if (f < nvar) {
#pragma unroll 9
        for (int64_t r = rg; r < nrow; r += 128) {
            const double x = A[r * lda + f];
            sa = sa + __ldg(wa + r) * x;
            sb = sb + __ldg(wb + r) * x;
        }
    }
    sa_s[rg][fi] = sa;
    sb_s[rg][fi] = sb;
    __syncthreads();
    if (threadIdx.x < 32) {                              // 8 runs of 16 row groups per field, then the 8 runs
        const int run = threadIdx.x >> 2;
        double ta = 0, tb = 0;
#pragma unroll
        for (int g = 0; g < 16; ++g) { ta = ta + sa_s[run * 16 + g][fi]; tb = tb + sb_s[run * 16 + g][fi]; }
        pa_s[run][fi] = ta;
        pb_s[run][fi] = tb;
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        double ta = 0, tb = 0;
#pragma unroll
        for (int g = 0; g < 8; ++g) { ta = ta + pa_s[g][fi]; tb = tb + pb_s[g][fi]; }
        fac_s[fi] = ta * (1. / tb);
    }
    __syncthreads();
    if (f < nvar) {
        const double fc = fac_s[fi];
#pragma unroll 9
        for (int64_t r = rg; r < nrow; r += 128) As[r * ldas + f] = A[r * lda + f] * fc;
    }
}
// v[c] = sum_r M[r,c] * wM[r], columns walked in CSC order (rows ascending): deterministic
__global__ void cons_v_kernel(const int32_t *colptr, const uint32_t *perm, const int32_t *row, const double *val, const double *wM,
                              int64_t ncol, double *v)
{
    int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncol) return;
    double s = 0;
    for (int32_t j = colptr[c]; j < colptr[c + 1]; ++j) { const uint32_t e = perm[j]; s += val[e] * wM[row[e]]; }
    v[c] = s;
}

static void compute_cons_v(Weighted &W, cudaStream_t s)
{
    if (W.cons_v.n == (size_t)W.ncol && W.ncol > 0) return;
    const int64_t nnz = W.nnz, ncol = W.ncol, nrow = W.nrow;
    W.cons_v.alloc((size_t)ncol, s);
    W.cons_v.zero();
    if (nnz == 0) return;
    DBuf<int32_t> row((size_t)nnz, s), csorted((size_t)nnz, s), colptr((size_t)ncol + 1, s);
    DBuf<uint32_t> k0((size_t)nnz, s), k1((size_t)nnz, s), p0((size_t)nnz, s), p1((size_t)nnz, s);
    expand_rows_kernel<<<(unsigned)cdiv(nrow, 128), 128, 0, s>>>(W.rowptr.p, nrow, row.p);
    IB_LAUNCHED();
    IB_CUDA(cudaMemcpyAsync(k0.p, W.col.p, nnz * sizeof(int32_t), cudaMemcpyDeviceToDevice, s));
    iota_u32_kernel<<<(unsigned)cdiv(nnz, 256), 256, 0, s>>>(p0.p, nnz);
    IB_LAUNCHED();
    const int w1 = radix_sort_pairs<uint32_t, uint32_t>(k0.p, p0.p, k1.p, p1.p, nnz, bits_for(ncol), s);
    u32_to_i32_copy_kernel<<<(unsigned)cdiv(nnz, 256), 256, 0, s>>>(w1 ? k1.p : k0.p, nnz, csorted.p);
    IB_LAUNCHED();
    colptr.zero();
    seg_ptr_kernel<<<(unsigned)cdiv(nnz, 256), 256, 0, s>>>(csorted.p, nnz, ncol, colptr.p);
    IB_LAUNCHED();
    cons_v_kernel<<<(unsigned)cdiv(ncol, 128), 128, 0, s>>>(colptr.p, w1 ? p1.p : p0.p, row.p, W.val.p, W.wM.p, ncol, W.cons_v.p);
    IB_LAUNCHED();
}

// Launch a fix-up kernel so that its launch overlaps the tail of the SpMM kernel before it (programmatic dependent
// launch); the kernel itself waits with griddepcontrol.wait before touching the partial sums.
template <class... KArgs, class... Args>
static void launch_dependent(void (*kern)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t s, Args... args)
{
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    IB_CUDA(cudaLaunchKernelEx(&cfg, kern, KArgs(args)...));
    IB_LAUNCHED();
}

// cudaFuncSetAttribute is per device: every device sets the dynamic shared-memory opt-ins on its first apply
struct DeviceAttrs { bool set = false; int max_smem = 0; };
static DeviceAttrs &device_attrs()
{
    static std::mutex mu;
    static DeviceAttrs da[IB_MAX_DEVICES];
    const int dev = device_index();
    std::lock_guard<std::mutex> lk(mu);
    DeviceAttrs &d = da[dev];
    if (!d.set) {
        IB_CUDA(cudaDeviceGetAttribute(&d.max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
        IB_CUDA(cudaFuncSetAttribute(spmm_xs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, d.max_smem));
        IB_CUDA(cudaFuncSetAttribute(spmm_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, d.max_smem));
        IB_CUDA(cudaFuncSetAttribute(spmm_gc_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gc_smem_bytes<4>()));
        IB_CUDA(cudaFuncSetAttribute(spmm_gc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gc_smem_bytes<2>()));
        IB_CUDA(cudaFuncSetAttribute(spmm_gc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gc_smem_bytes<1>()));
        d.set = true;
    }
    return d;
}

void apply_dev(Weighted &W, const double *dX, int64_t ldx, int nvar, double *dY, int64_t ldy, bool force_conservation,
               cudaStream_t s, int64_t row_begin, int64_t row_end)
{
    ensure_device();
    if (nvar <= 0 || W.nrow == 0) return;
    // row_end < 0: the whole matrix through the ordinary entry points.  A caller of the row-block entry point gets the
    // row-block kernels for every block, [0, nrow) included, so that blocks and their union agree bit for bit.
    const bool row_api = row_end >= 0;
    if (row_end < 0) row_end = W.nrow;
    const bool ranged = !(row_begin == 0 && row_end == W.nrow);
    if (ranged) {
        // row-block sharding (output rows split across GPUs, inputs replicated): groups of short rows never cross a
        // multiple of GROUP_ROWS, so blocks aligned to it own whole items
        if (row_begin < 0 || row_end > W.nrow || row_begin > row_end || row_begin % GROUP_ROWS || (row_end % GROUP_ROWS && row_end != W.nrow))
            fail(IB_ERR_ARG, "apply: row block [%ld, %ld) must be aligned to %d rows", (long)row_begin, (long)row_end, GROUP_ROWS);
        if (!W.conservative && force_conservation)
            fail(IB_ERR_ARG, "apply: the conservation correction needs all rows; apply row blocks with force_conservation=0");
        if (row_begin == row_end) return;
    }
    const int rlo = (int)row_begin, rhi = (int)row_end;
    // ---- conservation correction (ibmisc Weighted_Eigen::apply_e; matrix_formats.rst:156-165), fused as a prologue:
    //   B = B0 * (Mw.A)/(wM.B0)  with  wM.(M A) = (M^T wM).A =: v.A ,
    // so both dot products run over the (small) input fields, the factor is folded into a scaled copy of A, and the
    // large output is written once instead of being re-read twice.
    if (ldx < nvar || ldy < nvar) fail(IB_ERR_ARG, "apply: leading dimension smaller than nvar");
    DBuf<double> Xscaled;
    if (!W.conservative && force_conservation) {
        compute_cons_v(W, s);
        const int64_t ldxs = (nvar + 1) & ~1;
        Xscaled.alloc((size_t)W.ncol * ldxs, s);
        cons_scale_kernel<<<(unsigned)cdiv(nvar, 4), 512, 0, s>>>(W.Mw.p, W.cons_v.p, dX, ldx, W.ncol, nvar, Xscaled.p, ldxs);
        IB_LAUNCHED();
        dX = Xscaled.p;
        ldx = ldxs;
    }
    // wide matrices with about two nonzeros per column (EvI): one pass over X, no second kernel (apply_stream.cu)
    if (!row_api && g_tune.use_stream && apply_stream(W, nullptr, dX, ldx, nvar, dY, ldy, nullptr, 0, s)) return;
    make_plan(W, s);
    ApplyPlan &P = W.plan;
    const bool vec2 = (nvar % 2 == 0) && (ldx % 2 == 0) && (ldy % 2 == 0) && ((uintptr_t)dX % 16 == 0) && ((uintptr_t)dY % 16 == 0);
    const int64_t ldp = (nvar + 1) & ~1;
    DBuf<double> partials((size_t)P.npartials * ldp, s);
    DeviceAttrs &da = device_attrs();
    bool done = false;
    // distinct-column groups: every row has <= GC_S nonzeros (IvE, IvA, EvA)
    if (P.gc && vec2 && nvar >= 32 && P.gc_nitems * (int64_t)((nvar + 63) / 64) < (int64_t(1) << 31)) {
        int nq = nvar <= 64 ? 1 : 2;
        if (g_tune.gc_nq == 1 || g_tune.gc_nq == 2 || g_tune.gc_nq == 4) nq = g_tune.gc_nq;
        const int tw = 64 * nq;
        const int ntile = (nvar + tw - 1) / tw;
        GcArgs a{P.gc_items.p, P.gc_ucol.p, P.gc_recs.p, P.gc_nitems, ntile, dX, ldx, nvar, dY, ldy, rlo, rhi, P.gc_counter.p};
        const size_t smem = nq == 4 ? gc_smem_bytes<4>() : nq == 2 ? gc_smem_bytes<2>() : gc_smem_bytes<1>();
        auto kern = nq == 4 ? spmm_gc_kernel<4> : nq == 2 ? spmm_gc_kernel<2> : spmm_gc_kernel<1>;
        int occ = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, GC_WARPS * 32, smem);
        occ = std::max(occ, 1);
        const int64_t nblk = std::max<int64_t>(1, std::min<int64_t>(cdiv(a.nitems * ntile, GC_WARPS), (int64_t)sm_count() * occ));
        if (a.nitems > 0) {
            kern<<<(unsigned)nblk, GC_WARPS * 32, smem, s>>>(a);
            IB_LAUNCHED();
        }
        return;
    }
    // tall matrices with longer rows that share columns (sigma-smoothed IvE / IvA): dense 8 x 4 blocks of 32-row items on the fp64 tensor-core path
    if (g_tune.use_group && !P.gc && P.max_row_len <= 128 && apply_group(W, dX, ldx, nvar, dY, ldy, rlo, rhi, s)) return;
    // X-stationary path: tall matrix, few columns, no long rows
    if (vec2 && !ranged && W.ncol > 0 && W.nrow >= 8 * W.ncol && P.max_row_len <= 256 && nvar >= 8 && 2 * W.nnz >= 3 * W.nrow) {
        int F = 64;
        while (F > nvar) F >>= 1;
        while (F >= 8 && (size_t)W.ncol * F * 8 > (size_t)da.max_smem - 1024) F >>= 1;
        if (F >= 8) {
            const size_t smem = (size_t)W.ncol * F * 8;
            int occ = 0;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, spmm_xs_kernel, XS_THREADS, smem);
            occ = std::max(1, std::min(occ, 4));
            XsArgs a{W.rowptr.p, W.col.p, W.val.p, dX, ldx, dY, ldy, W.nrow, (int)W.ncol, nvar, F, (nvar + F - 1) / F, 0, 0, 2048};
            a.nchunk = cdiv(W.nrow, a.chunk_rows);
            a.nunits = a.nchunk * a.nslice;
            const int64_t nblk = std::max<int64_t>(1, std::min<int64_t>(a.nunits, (int64_t)sm_count() * occ));
            spmm_xs_kernel<<<(unsigned)nblk, XS_THREADS, smem, s>>>(a);
            IB_LAUNCHED();
            return;            // rows are never split: nothing to fix up
        }
    }
    if (vec2 && nvar >= 64) {
        const int tw = std::min(nvar, TMA_MAX_TW);
        const int ntile = (nvar + tw - 1) / tw;
        const int64_t nunits = P.nitems * ntile;
        const int bps = std::max(1, g_tune.tma_bps);
        const size_t stagebytes = (size_t)tw * 8 * TMA_ROWS + TMA_HDR;
        const size_t budget = (size_t)(da.max_smem + 1024) / bps - 1024 - 256;
        const int nstage = (int)std::min<size_t>(15, budget / stagebytes);
        if (nstage >= 2) {
            const size_t smem = 256 + (size_t)nstage * stagebytes;
            SpmmArgs a{P.items.p, nunits, ntile, tw, nstage, W.rowptr.p, W.col.p, W.val.p, dX, ldx, nvar, dY, ldy, partials.p, ldp, rlo, rhi};
            const int64_t slots = (int64_t)sm_count() * bps;
            const int64_t per = std::max<int64_t>(1, cdiv(nunits, slots));
            const int64_t nblk = std::max<int64_t>(1, cdiv(nunits, per));
            spmm_tma_kernel<<<(unsigned)nblk, TMA_CONSUMERS + 32, smem, s>>>(a);
            IB_LAUNCHED();
            done = true;
        }
    }
    if (!done) {
        const int VEC = vec2 ? 2 : 1;
        const int lanes_needed = (nvar + VEC - 1) / VEC;
        int L = 1;
        while (L < 32 && L < lanes_needed) L <<= 1;
        const int ntile = (lanes_needed + L - 1) / L;
        const int64_t nunits = P.nitems * ntile;
        const int G = 32 / L;
        int64_t blocks = std::min<int64_t>(cdiv(cdiv(nunits, G), 8), (int64_t)sm_count() * 8);
        if (blocks < 1) blocks = 1;
        if (VEC == 2)
            spmm_ldg_kernel<2><<<(unsigned)blocks, 256, 0, s>>>(P.items.p, nunits, ntile, L, W.rowptr.p, W.col.p, W.val.p, dX, ldx, nvar, dY, ldy, partials.p, ldp, rlo, rhi);
        else
            spmm_ldg_kernel<1><<<(unsigned)blocks, 256, 0, s>>>(P.items.p, nunits, ntile, L, W.rowptr.p, W.col.p, W.val.p, dX, ldx, nvar, dY, ldy, partials.p, ldp, rlo, rhi);
        IB_LAUNCHED();
    }
    if (P.nsplit_rows > 0) {
        if (vec2) {
            int64_t tot = P.nsplit_rows * (nvar / 2);
            launch_dependent(spmm_fixup_kernel<2>, (unsigned)cdiv(tot, 256), 256, 0, s, P.split_rows.p, P.nsplit_rows, partials.p, ldp, nvar, dY, ldy, rlo, rhi);
        } else {
            int64_t tot = P.nsplit_rows * (int64_t)nvar;
            launch_dependent(spmm_fixup_kernel<1>, (unsigned)cdiv(tot, 256), 256, 0, s, P.split_rows.p, P.nsplit_rows, partials.p, ldp, nvar, dY, ldy, rlo, rhi);
        }
    }
}

}   // namespace ib
