// apply_group.cu -- Y = M X for TALL matrices whose neighbouring rows name mostly the same columns: the sigma-smoothed
// IvE / IvA (RegridMatrices_Dynamic.cpp:236-248, smoother.cpp:43-99): ~16 nonzeros per row, and the ~36 distinct columns
// of 32 consecutive ice cells are the elevation classes their common neighbourhood feeds.
//
// Why not the other kernels.  X is small (E or A cells: L2-resident), Y is large (ice cells).  A kernel that reads one X
// row per nonzero moves 16x the output through L2 (0.4 ms, what spmm_xs_kernel measured); keeping X rows in shared memory
// and multiplying with DFMA still costs one shared-memory read per column and 36 LDS.128 per 128 DFMA -- measured 182 us
// with the fp64 pipe 27 % busy, warps waiting on shared-memory loads (ncu: short scoreboard).  So the product is done as
// small DENSE blocks on the fp64 tensor-core path (DMMA.8x8x4; B200 measured 72.5 G mma/s = 37 TFLOP/s, the same flops as
// DFMA but 8 operand loads per ~10 mma):
//   item  = 32 consecutive rows = four m-tiles of 8 rows; its distinct columns, ordered by (first, last m-tile that names
//           them) so that most 8-row x 4-column blocks of the item are either dense or empty (3.3 M DMMA at 5 km against
//           5.1 M for the dense items), cut into chunks of 8 columns;
//   chunk = the X rows of its columns (ONE 1-D bulk async copy each: 4 KB at 512 fields) + the coefficients of its
//           2 k-steps x 4 m-tiles laid out in DMMA A-fragment order, a bit per non-empty block, a bit per m-tile that
//           ends here, and a 32-bit row-presence mask per column; prepared once per matrix; a chunk is one stage of a
//           shared-memory ring fed by a producer warp;
//   a consumer warp owns a 32-field strip of the item: 4 x 4 accumulator tiles; per chunk 8 B-fragment loads (X), one
//   A-fragment load and 4 DMMA per non-empty block, selected by a switch over the 16 block masks of a k-step (a
//   predicated-off DMMA still occupies the pipe).  X rows sit 32 bytes apart modulo 128 in shared memory so that the
//   B-fragment loads (4 rows x 8 fields) are conflict-free.  An m-tile is written as soon as its last chunk is done.
// Rows that do not name a column of a non-empty block carry a zero coefficient, which adds nothing while x is finite;
// where a warp sees a NaN / Inf in the chunk's X values the presence bits take over (scalar FMA path), so a NaN in X
// reaches only the rows that name its cell (structural zeros are never multiplied into a sum).  The terms of a row are
// added four columns at a time inside the tensor core, so results agree with the oracle to rounding (1e-12 gate), not bit
// for bit.
//
// Roofline: HBM (Y written once: 12*nnz + 8*nvar*(nrow + ncol) algorithmic bytes).  What a consumer warp's time goes to
// at 5 km / 512 fields (tools/group_trace.py, 77 us per CTA): DMMA issue on a saturated pipe 45 us (= the 3.3 M DMMA at
// the measured rate), the loads / checks / addressing ahead of them 19 us (issue-bound: the four warps of a scheduler
// move in step, so this does not overlap the other warps' DMMA), stores 7 us, waiting for a stage 6 us.  The CTAs get
// equal shares of a fitted cost (item, chunk, block), 72 .. 85 us; the kernel takes 84 us = 0.54 of the HBM roofline.
// History: every warp its own pipeline with 512-byte copies 380 us -> CTA ring, 16-row items, DFMA 182 us -> DMMA 139 us
// -> A fragments before the first DMMA + 112 registers for the consumers (setmaxnreg) 105 us -> switch instead of
// predicates 102 us -> CTA shares read from a table instead of a binary search at the head of every CTA 99 us -> shares
// by fitted cost 87 us -> programmatic dependent launch 84 us.
#include <algorithm>
#include <mutex>

#include "scan.cuh"
#include "tma.cuh"
#include "types.cuh"

namespace ib {

constexpr int GR_R = 32;                 // rows per item (four 8-row m-tiles)
constexpr int GR_CH = 8;                 // columns per chunk (two k-steps of 4)
constexpr int GR_MAXD = 128;             // distinct columns per item the plan accepts
constexpr int64_t GR_MAX_ITEMS = 3500000; // an item costs at most (1140 + 16 * (491 + 8 * 69)) / 16 < 1200 units: sum < 2^32
// chunk record: ncols, block mask (bit 4*kstep + mtile), done mask (bit mtile), pad | cols[8] | rowmask[8] | coef[kstep 2][mtile 4][lane 32]
constexpr int GR_OFF_COLS = 16, GR_OFF_RMASK = GR_OFF_COLS + 4 * GR_CH, GR_OFF_COEF = GR_OFF_RMASK + 4 * GR_CH;
constexpr int GR_REC = GR_OFF_COEF + 8 * GR_CH * GR_R;
constexpr int GR_SKEW = 32;              // bytes added to the shared-memory stride of an X row
static_assert(GR_OFF_COEF % 16 == 0 && GR_REC % 16 == 0, "bulk copies move multiples of 16 bytes");

// ---- plan: one warp per item, lane = row.  The rows' columns ascend (CSR), so the item's distinct columns come out of
// a 32-way merge: the smallest head over the lanes is the next column, the lanes holding it are its rows.
__device__ __forceinline__ int gr_next_column(const int32_t *col, int32_t &j, int32_t je, uint32_t &rows, bool &sorted)
{
    const int head = j < je ? col[j] : INT_MAX;
    const int m = __reduce_min_sync(0xffffffffu, head);
    rows = __ballot_sync(0xffffffffu, head == m);
    if (head == m)
        while (j < je && col[j] == m) ++j;                 // duplicate entries of a caller-supplied CSR
    if (j < je && col[j] < m) sorted = false;
    return m;
}

__global__ void __launch_bounds__(128) gr_count_kernel(const int32_t *__restrict__ rowptr, const int32_t *__restrict__ col, int64_t nrow,
                                                       int64_t nitems, uint32_t *__restrict__ counts, int *__restrict__ refuse)
{
    const int64_t it = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (it >= nitems) return;
    const int lane = threadIdx.x & 31;
    const int64_t r = it * GR_R + lane;
    int32_t j = r < nrow ? rowptr[r] : 0;
    const int32_t je = r < nrow ? rowptr[r + 1] : 0;
    bool sorted = true;
    int D = 0;
    for (;; ++D) {
        uint32_t rows;
        if (gr_next_column(col, j, je, rows, sorted) == INT_MAX) break;
    }
    if (!__all_sync(0xffffffffu, sorted) || D > GR_MAXD) { if (lane == 0) atomicExch(refuse, 1); D = 0; }
    if (lane == 0) counts[it] = (uint32_t)max(1, (D + GR_CH - 1) / GR_CH);
}

struct GrCountLoad {
    const uint32_t *counts;
    __device__ uint32_t operator()(int64_t it) const { return counts[it]; }
};
struct GrItemEmit {
    int4 *items;
    int64_t nrow;
    __device__ void operator()(int64_t it, uint32_t prefix, uint32_t nch) const
    {
        const int64_t r0 = it * GR_R;
        items[it] = make_int4((int)r0, (int)prefix, (int)(prefix + nch), (int)min((int64_t)GR_R, nrow - r0));
    }
};

struct GrPrefixEmit {
    uint32_t *prefix;
    __device__ void operator()(int64_t it, uint32_t before, uint32_t) const { prefix[it] = before; }
};

__global__ void __launch_bounds__(128) gr_emit_kernel(const int32_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                                                      const double *__restrict__ val, int64_t nrow, int64_t nitems,
                                                      const int4 *__restrict__ items, unsigned char *__restrict__ chunks,
                                                      uint32_t *__restrict__ weight)
{
    __shared__ int32_t s_cols[4][GR_MAXD];
    __shared__ uint32_t s_rows[4][GR_MAXD];
    __shared__ unsigned char s_pos[4][GR_MAXD];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t it = (int64_t)blockIdx.x * 4 + w;
    if (it >= nitems) return;
    const int64_t r = it * GR_R + lane;
    const int32_t jb = r < nrow ? rowptr[r] : 0, je = r < nrow ? rowptr[r + 1] : 0;
    int32_t j = jb;
    bool sorted = true;
    int D = 0;
    for (;; ++D) {
        uint32_t rows;
        const int c = gr_next_column(col, j, je, rows, sorted);
        if (c == INT_MAX) break;
        if (lane == 0 && D < GR_MAXD) { s_cols[w][D] = c; s_rows[w][D] = rows; }
    }
    if (D > GR_MAXD) D = 0;                                  // refused by the count pass; keep the writes in bounds
    __syncwarp();
    // order the columns by (first, last) m-tile that names them, then by column: a stable counting sort over 16 keys
    int key[GR_MAXD / 32];
#pragma unroll
    for (int b = 0; b < GR_MAXD / 32; ++b) {
        const int d = lane + 32 * b;
        key[b] = -1;
        if (d < D) {
            const uint32_t m = s_rows[w][d];
            const uint32_t tm = ((m & 0xffu) != 0) | (((m >> 8) & 0xffu) != 0) << 1 | (((m >> 16) & 0xffu) != 0) << 2 | ((m >> 24) != 0) << 3;
            key[b] = (__ffs(tm) - 1) * 4 + (31 - __clz(tm));
        }
    }
    int base = 0;
    for (int k = 0; k < 16; ++k)
#pragma unroll
        for (int b = 0; b < GR_MAXD / 32; ++b) {
            const uint32_t bal = __ballot_sync(0xffffffffu, key[b] == k);
            if (key[b] == k) s_pos[w][lane + 32 * b] = (unsigned char)(base + __popc(bal & ((1u << lane) - 1u)));
            base += __popc(bal);
        }
    __syncwarp();
    const int4 item = items[it];
    const int nch = item.z - item.y;
    unsigned char *rec0 = chunks + (size_t)item.y * GR_REC;
    // headers and zeroed tables
    for (int q = 0; q < nch; ++q) {
        unsigned char *rec = rec0 + (size_t)q * GR_REC;
        int32_t *h = reinterpret_cast<int32_t *>(rec);
        // ncols counts whole k-steps: the slots past the last column repeat it with zero coefficients and an empty row
        // mask, so that the kernel never multiplies stale shared memory
        const int ncq = max(0, min(GR_CH, D - q * GR_CH));
        if (lane < 4) h[lane] = lane == 0 ? (ncq + 3) / 4 * 4 : 0;
        if (lane < 2 * GR_CH) h[4 + lane] = (lane < GR_CH && lane >= ncq && ncq > 0) ? s_cols[w][0] : 0;   // cols and row masks
        double *cf = reinterpret_cast<double *>(rec + GR_OFF_COEF);
        for (int k = lane; k < GR_CH * GR_R; k += 32) cf[k] = 0.0;
    }
    __syncwarp();
    for (int d = lane; d < D; d += 32) {
        const int p = s_pos[w][d];
        unsigned char *rec = rec0 + (size_t)(p / GR_CH) * GR_REC;
        const int k = p % GR_CH;
        const uint32_t m = s_rows[w][d];
        reinterpret_cast<int32_t *>(rec + GR_OFF_COLS)[k] = s_cols[w][d];
        reinterpret_cast<uint32_t *>(rec + GR_OFF_RMASK)[k] = m;
        const uint32_t tm = ((m & 0xffu) != 0) | (((m >> 8) & 0xffu) != 0) << 1 | (((m >> 16) & 0xffu) != 0) << 2 | ((m >> 24) != 0) << 3;
        atomicOr(reinterpret_cast<int32_t *>(rec) + 1, (int)(tm << (4 * (k / 4))));
    }
    // header word 2: the m-tiles whose last column lies in this chunk (their rows are complete and can be written)
    {
        int last[4] = {-1, -1, -1, -1};
        for (int d = lane; d < D; d += 32) {
            const uint32_t m = s_rows[w][d];
            const int p = s_pos[w][d];
#pragma unroll
            for (int t = 0; t < 4; ++t)
                if ((m >> (8 * t)) & 0xffu) last[t] = max(last[t], p);
        }
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int lt = __reduce_max_sync(0xffffffffu, last[t]);
            if (lane == 0) atomicOr(reinterpret_cast<int32_t *>(rec0 + (size_t)(lt < 0 ? nch - 1 : lt / GR_CH) * GR_REC) + 2, 1 << t);
        }
    }
    // estimated cost of the item, in ns: least-squares fit of the CTAs' lifetimes at 512 fields (tools/group_trace.py)
    // against their items, chunks and non-empty blocks: 1140 per item (stores, accumulator reset), 491 per chunk (loads,
    // checks, barriers), 69 per block (4 DMMA on a pipe shared by four warps); residual 1.4 us per CTA
    {
        __syncwarp();
        uint32_t wq = 0;
        for (int q = lane; q < nch; q += 32)
            wq += 491u + 69u * (uint32_t)__popc(reinterpret_cast<const int32_t *>(rec0 + (size_t)q * GR_REC)[1]);
        wq = __reduce_add_sync(0xffffffffu, wq);
        if (lane == 0) weight[it] = (wq + 1140u + 8u) >> 4;      // units of 16 ns: the running sum stays below 2^32 (GR_MAX_ITEMS)
    }
    // coefficients: lane = row; A-fragment order: [k-step][m-tile][(row % 8) * 4 + column % 4]
    for (int32_t e = jb; e < je; ++e) {
        const int32_t c = col[e];
        int lo = 0, hi = D;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (s_cols[w][mid] < c) lo = mid + 1; else hi = mid; }
        if (lo >= D) continue;                               // only for a refused item
        const int p = s_pos[w][lo];
        double *cf = reinterpret_cast<double *>(rec0 + (size_t)(p / GR_CH) * GR_REC + GR_OFF_COEF);
        const int k = p % GR_CH;
        // duplicate (row, column) entries add up, in storage order (a row belongs to one lane)
        cf[(k / 4) * (4 * 32) + (lane / 8) * 32 + (lane % 8) * 4 + (k % 4)] += val[e];
    }
}

static bool make_group_plan(Weighted &W, cudaStream_t s)
{
    GroupPlan &P = W.group_plan;
    if (P.tried) return P.ok;
    P.tried = true;
    if (W.nrow < 4 * GR_R || W.nnz < 4 * W.nrow || W.nrow < 4 * W.ncol || W.nnz >= (int64_t(1) << 31)) return false;
    const int64_t nitems = cdiv(W.nrow, GR_R);
    if (nitems > GR_MAX_ITEMS) return false;
    DBuf<int> refuse(1, s);
    refuse.zero();
    DBuf<uint32_t> counts((size_t)nitems, s);
    gr_count_kernel<<<(unsigned)cdiv(nitems, 4), 128, 0, s>>>(W.rowptr.p, W.col.p, W.nrow, nitems, counts.p, refuse.p);
    IB_LAUNCHED();
    ScanState<uint32_t> st;
    GrCountLoad L{counts.p};
    transform_scan<uint32_t>(st, nitems, L, AddU32{}, 0u, s);
    uint32_t nch = 0;
    int hrefuse = 0;
    IB_CUDA(cudaMemcpyAsync(&nch, st.total.p, sizeof nch, cudaMemcpyDeviceToHost, s));
    refuse.download(&hrefuse, 1);
    IB_CUDA(cudaStreamSynchronize(s));
    // too many distinct columns per 32 rows (or unsorted rows), or tables much larger than the matrix: the rows do not
    // share columns
    if (hrefuse || (int64_t)nch * GR_REC > 6 * 12 * W.nnz + (64 << 10)) return false;
    P.nitems = nitems;
    P.nchunks = nch;
    P.items.alloc((size_t)nitems, s);
    P.chunks.alloc((size_t)nch * GR_REC, s);
    scan_apply<uint32_t>(st, L, AddU32{}, 0u, GrItemEmit{P.items.p, W.nrow}, s);
    gr_emit_kernel<<<(unsigned)cdiv(nitems, 4), 128, 0, s>>>(W.rowptr.p, W.col.p, W.val.p, W.nrow, nitems, P.items.p, P.chunks.p, counts.p);
    IB_LAUNCHED();
    // counts now holds the items' estimated costs; their running sum is what the CTAs split evenly
    P.wprefix.alloc((size_t)nitems + 1, s);
    ScanState<uint32_t> sw;
    transform_scan<uint32_t>(sw, nitems, L, AddU32{}, 0u, s);
    scan_apply<uint32_t>(sw, L, AddU32{}, 0u, GrPrefixEmit{P.wprefix.p}, s);
    IB_CUDA(cudaMemcpyAsync(P.wprefix.p + nitems, sw.total.p, sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
    IB_CUDA(cudaStreamSynchronize(s));
    P.ok = true;
    return true;
}

// CTA b of `grid` owns the items of [item_lo, item_hi) that start in its share of the block's estimated cost (done once
// per row block: twenty dependent loads at the head of every CTA cost 15 us per launch when the kernel searched itself)
__global__ void gr_split_kernel(const uint32_t *__restrict__ wprefix, int64_t item_lo, int64_t item_hi, int grid, int *__restrict__ split)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b > grid) return;
    const uint64_t w_lo = wprefix[item_lo], w_hi = wprefix[item_hi];
    const uint64_t w = w_lo + (w_hi - w_lo) * (uint64_t)b / (uint64_t)grid;
    int64_t lo = item_lo, hi = item_hi;                      // first item that starts at or after w
    while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (wprefix[mid] < w) lo = mid + 1; else hi = mid; }
    split[b] = b == grid ? (int)item_hi : (int)lo;
}

// ---- kernel: the ring of spmm_stream_kernel -- a producer warp feeds stages (= chunks) to up to 16 consumer warps
struct GrArgs {
    const int4 *items;
    const unsigned char *chunks;
    const int *split;              // [gridDim.x + 1] first item of every CTA
    int ntile, tw, nstage;
    const double *X;
    int64_t ldx;
    int nvar;
    double *Y;
    int64_t ldy;
};

#ifdef IB_GROUP_TRACE       // tools/build_variants.py: per-warp cycle counts of the last launch (never in the product library)
__device__ unsigned long long g_group_trace[256 * 17 * 8];      // warp 16 = producer
#define GTR_DECL unsigned long long tr_wait = 0, tr_math = 0, tr_store = 0, tr_pre = 0, tr_g0; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(tr_g0)); const long long tr_t0 = clock64(); long long tr_t = tr_t0
#define GTR_ADD(acc) do { const long long now_ = clock64(); acc += now_ - tr_t; tr_t = now_; } while (0)
#define GTR_FLUSH() do { if (lane == 0 && blockIdx.x < 256) { unsigned long long *o_ = g_group_trace + ((size_t)blockIdx.x * 17 + warp) * 8; o_[4] = tr_pre; { unsigned long long g1_; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g1_)); o_[5] = g1_ - tr_g0; o_[6] = tr_g0; o_[7] = g1_; } \
    o_[0] = tr_wait; o_[1] = tr_math; o_[2] = tr_store; o_[3] = clock64() - tr_t0; } } while (0)
#else
#define GTR_DECL do { } while (0)
#define GTR_ADD(acc) do { } while (0)
#define GTR_FLUSH() do { } while (0)
#endif

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b)
{
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
// The DMMAs of the non-empty blocks (bit t of M: m-tile t) of one k-step.  Selected by a switch over the 16 masks rather
// than a predicate per block: a predicated-off DMMA still occupies the pipe (ncu: the pipe was busy for the DENSE count).
template <int M>
__device__ __forceinline__ void gr_blocks(double (&acc)[4][4][2], const double (&av)[4], const double (&b)[4])
{
#pragma unroll
    for (int t = 0; t < 4; ++t)
        if (M >> t & 1) {
#pragma unroll
            for (int j = 0; j < 4; ++j) dmma884(acc[t][j][0], acc[t][j][1], av[t], b[j]);
        }
}
__device__ __forceinline__ bool gr_finite(double x) { return (__double2hiint(x) & 0x7ff00000) != 0x7ff00000; }

// Launched with GR_THREADS threads: warps 0..15 consume (those past ncw retire at once), warp 16 produces, warps 17..19
// only make its warpgroup whole.  Seventeen working warps would cap every thread at 96 registers (five warps on one
// scheduler); the producer's warpgroup hands its registers back (setmaxnreg) so that the consumers, which hold 64
// accumulator registers each, run with 112.
constexpr int GR_THREADS = 640;
__global__ void __launch_bounds__(GR_THREADS, 1) smooth_group_kernel(const __grid_constant__ GrArgs a)
{
    extern __shared__ __align__(128) unsigned char gr_raw[];
    uint64_t *full = reinterpret_cast<uint64_t *>(gr_raw);
    uint64_t *empty = full + a.nstage;
    unsigned char *data = gr_raw + 128;
    const uint32_t rowstride = (uint32_t)a.tw * 8u + GR_SKEW;
    const uint32_t stagebytes = rowstride * GR_CH + GR_REC;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ncw = a.tw / 32;
    if (threadIdx.x == 0) {
        for (int s = 0; s < a.nstage; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, ncw); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // a contiguous share of the items, balanced by their chunks (gr_split_kernel; a CTA walks neighbouring items: their
    // X rows overlap in L2)
    const int64_t i0 = __ldg(a.split + blockIdx.x), i1 = __ldg(a.split + blockIdx.x + 1);
    int stage = 0;
    uint32_t phase = 0;
    if (warp >= 16) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
        if (warp > 16 || i0 >= i1) return;
        pdl_wait();                      // X may be the previous kernel's output (the corrected copy of the input)
        // ===== producer: one stage per chunk; lane k < ncols copies the X row of column k, lane 31 the chunk record.
        // The column ids of a stage are loaded one stage ahead, so the only thing the producer ever waits for is a
        // free stage.
        GTR_DECL;
        for (int tile = 0; tile < a.ntile; ++tile) {
            const int f0 = tile * a.tw;
            const uint32_t tb = (uint32_t)min(a.tw, a.nvar - f0) * 8u;
            const int q0 = __ldg(a.items + i0).y, q1 = __ldg(a.items + i1 - 1).z;
            // metadata of FOUR chunks per load (lane = 8 * chunk + column slot), fetched a group ahead: the producer's
            // own loads were what paced it when it looked one chunk ahead (a dependent L2 round trip per stage)
            const int jq = lane >> 3, k = lane & 7;
            auto meta = [&](int qg, int &nc, int32_t &c) {
                nc = 0; c = 0;
                if (qg + jq < q1) {
                    const unsigned char *rec = a.chunks + (size_t)(qg + jq) * GR_REC;
                    nc = __ldg(reinterpret_cast<const int32_t *>(rec));
                    c = __ldg(reinterpret_cast<const int32_t *>(rec + GR_OFF_COLS) + k);
                }
            };
            int nc4;
            int32_t c;
            meta(q0, nc4, c);
            for (int qg = q0; qg < q1; qg += 4) {
                int nc4n;
                int32_t cn;
                meta(qg + 4, nc4n, cn);
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const int q = qg + jj;
                    if (q >= q1) break;
                    const int nc = __shfl_sync(0xffffffffu, nc4, 8 * jj);
                    unsigned char *sp = data + (size_t)stage * stagebytes;
                    GTR_ADD(tr_math);
                    if (lane == 0) {
                        mbar_wait(empty + stage, phase ^ 1u);
                        mbar_expect_tx(full + stage, tb * (uint32_t)nc + GR_REC);
                    }
                    __syncwarp();
                    GTR_ADD(tr_wait);
                    if (jq == jj && k < nc) bulk_g2s(sp + (size_t)k * rowstride, a.X + (int64_t)c * a.ldx + f0, tb, full + stage);
                    if (lane == 8 * ((jj + 1) & 3)) bulk_g2s(sp + (size_t)GR_CH * rowstride, a.chunks + (size_t)q * GR_REC, GR_REC, full + stage);
                    if (++stage == a.nstage) { stage = 0; phase ^= 1u; }
                }
                nc4 = nc4n; c = cn;
            }
        }
        GTR_FLUSH();
        return;
    }
    asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
    if (warp >= ncw || i0 >= i1) return;
    // ===== consumers: warp w owns fields f0 + 32 w .. + 31 of all 32 rows of the item: acc[m-tile][n-tile] in DMMA
    // C-fragment layout (row 8 t + lane / 4, fields 8 j + 2 (lane % 4), + 1)
    const int fo = 32 * warp;
    const int g = lane >> 2, q4 = lane & 3;
    double acc[4][4][2];
    GTR_DECL;
    for (int tile = 0; tile < a.ntile; ++tile) {
        const int f0 = tile * a.tw;
        const int nf = min(a.tw, a.nvar - f0);               // fields of this tile
        const bool fields_full = fo + 32 <= nf;
        for (int64_t i = i0; i < i1; ++i) {
            const int4 it = __ldg(a.items + i);
            double *y = a.Y + (int64_t)it.x * a.ldy + f0 + fo + 2 * q4;
#pragma unroll
            for (int t = 0; t < 4; ++t)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[t][j][0] = acc[t][j][1] = 0.0;
            for (int q = it.y; q < it.z; ++q) {
                GTR_ADD(tr_math);
                mbar_wait(full + stage, phase);
                GTR_ADD(tr_wait);
                const unsigned char *sb = data + (size_t)stage * stagebytes;
                const unsigned char *rec = sb + (size_t)GR_CH * rowstride;
                // header: ncols (whole k-steps), block mask, done mask -- one 16-byte load
                const int4 hdr = *reinterpret_cast<const int4 *>(rec);
#ifdef IB_GROUP_SKIP_MATH       // tools/: delivery rate of the ring alone (results are wrong)
                const int nc = 0;
#else
                const int nc = hdr.x;
#endif
                const uint32_t bm = (uint32_t)hdr.y;
                const double *cf = reinterpret_cast<const double *>(rec + GR_OFF_COEF);
                if (nc > 0) {
                    // all fragments of the chunk before the first DMMA.  B: X[column 4 s + lane % 4][field fo + 8 j +
                    // lane / 4]; A: the four m-tiles of each k-step (empty blocks hold zeros).  The second k-step's X
                    // rows exist only if the chunk has more than four columns.
                    const unsigned char *xb = sb + (size_t)q4 * rowstride + (size_t)(fo + g) * 8;
                    const bool two = nc > 4;
                    double b[2][4], av[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        b[0][j] = *reinterpret_cast<const double *>(xb + 64 * j);
                        b[1][j] = two ? *reinterpret_cast<const double *>(xb + 4 * (size_t)rowstride + 64 * j) : 0.0;
                    }
#pragma unroll
                    for (int t = 0; t < 4; ++t) av[t] = cf[t * 32 + lane];
                    if (!fields_full) {
                        // fields past nvar were not copied: stale shared memory
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if (fo + 8 * j + g >= nf) b[0][j] = b[1][j] = 0.0;
                    }
#ifdef IB_GROUP_NOCHECK         // tools/: cost of the NaN / Inf check (results are wrong for non-finite input)
                    const bool fin = true;
#else
                    bool fin = true;
#pragma unroll
                    for (int j = 0; j < 4; ++j) fin = fin && gr_finite(b[0][j]) && gr_finite(b[1][j]);
#endif
                    const bool allfin = __all_sync(0xffffffffu, fin);
                    GTR_ADD(tr_pre);
                    if (allfin) {
#pragma unroll
                        for (int s = 0; s < 2; ++s) {
                            // the second k-step's A fragments take the registers of the first's once its DMMAs are issued
                            if (s == 1) {
#pragma unroll
                                for (int t = 0; t < 4; ++t) av[t] = cf[(4 + t) * 32 + lane];
                            }
                            switch (bm >> (4 * s) & 15u) {
                            case 1: gr_blocks<1>(acc, av, b[s]); break;
                            case 2: gr_blocks<2>(acc, av, b[s]); break;
                            case 3: gr_blocks<3>(acc, av, b[s]); break;
                            case 4: gr_blocks<4>(acc, av, b[s]); break;
                            case 5: gr_blocks<5>(acc, av, b[s]); break;
                            case 6: gr_blocks<6>(acc, av, b[s]); break;
                            case 7: gr_blocks<7>(acc, av, b[s]); break;
                            case 8: gr_blocks<8>(acc, av, b[s]); break;
                            case 9: gr_blocks<9>(acc, av, b[s]); break;
                            case 10: gr_blocks<10>(acc, av, b[s]); break;
                            case 11: gr_blocks<11>(acc, av, b[s]); break;
                            case 12: gr_blocks<12>(acc, av, b[s]); break;
                            case 13: gr_blocks<13>(acc, av, b[s]); break;
                            case 14: gr_blocks<14>(acc, av, b[s]); break;
                            case 15: gr_blocks<15>(acc, av, b[s]); break;
                            default: break;
                            }
                        }
                    } else {
                        // a NaN / Inf among the X values: only the rows that name a column may see it
#pragma unroll 1
                        for (int c = 0; c < nc; ++c) {
                            const uint32_t rm = *reinterpret_cast<const uint32_t *>(rec + GR_OFF_RMASK + 4 * c);
                            const unsigned char *xr = sb + (size_t)c * rowstride + (size_t)(fo + 2 * q4) * 8;
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                double2 x = *reinterpret_cast<const double2 *>(xr + 64 * j);
                                if (fo + 8 * j + 2 * q4 >= nf) x = make_double2(0.0, 0.0);
#pragma unroll
                                for (int t = 0; t < 4; ++t)
                                    if (rm >> (8 * t + g) & 1u) {
                                        const double v = cf[((c >> 2) * 4 + t) * 32 + g * 4 + (c & 3)];
                                        acc[t][j][0] = fma(v, x.x, acc[t][j][0]);
                                        acc[t][j][1] = fma(v, x.y, acc[t][j][1]);
                                    }
                            }
                        }
                    }
                }
                const uint32_t done = (uint32_t)hdr.z;
                __syncwarp();
                if (lane == 0) mbar_arrive(empty + stage);
                if (++stage == a.nstage) { stage = 0; phase ^= 1u; }
                // the m-tiles that end with this chunk are written now (8 rows x 256 bytes per warp), spreading the
                // output over the item instead of one burst of 16 stores per warp at its end; no later block of the
                // item touches their accumulators, which are cleared with all the others when the next item starts
                if (done) {
                    GTR_ADD(tr_math);
#pragma unroll
                    for (int t = 0; t < 4; ++t)
                        if (done >> t & 1u) {
#ifdef IB_GROUP_SKIP_STORE      // tools/: the ring without its output (results are wrong)
                            if (acc[0][0][0] == 12345.678)
#endif
                            if (8 * t + g < it.w) {
#pragma unroll
                                for (int j = 0; j < 4; ++j)
                                    if (fo + 8 * j + 2 * q4 < nf)
                                        *reinterpret_cast<double2 *>(y + (int64_t)(8 * t + g) * a.ldy + 8 * j) = make_double2(acc[t][j][0], acc[t][j][1]);
                            }
                        }
                    GTR_ADD(tr_store);
                }
            }
        }
    }
    GTR_FLUSH();
}

#ifdef IB_GROUP_TRACE
extern "C" int ib_debug_group_trace(unsigned long long *out, int nblk)
{
    cudaDeviceSynchronize();
    return (int)cudaMemcpyFromSymbol(out, g_group_trace, sizeof(unsigned long long) * 17 * 8 * nblk);
}
#endif

// Applies W through the group kernel; false (nothing launched) if the matrix does not fit it.
bool apply_group(Weighted &W, const double *dX, int64_t ldx, int nvar, double *dY, int64_t ldy, int row_lo, int row_hi, cudaStream_t s)
{
    if (nvar < 32 || nvar % 2 || ldx % 2 || ldy % 2 || (uintptr_t)dX % 16 || (uintptr_t)dY % 16) return false;
    if (row_lo % GR_R) return false;
    if (!make_group_plan(W, s)) return false;
    GroupPlan &P = W.group_plan;
    static int max_smem[IB_MAX_DEVICES] = {};
    const int dev = device_index();
    {
        static std::mutex mu;
        std::lock_guard<std::mutex> lk(mu);
        if (!max_smem[dev]) {
            int ms = 0;
            IB_CUDA(cudaDeviceGetAttribute(&ms, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
            IB_CUDA(cudaFuncSetAttribute(smooth_group_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ms));
            max_smem[dev] = ms;
        }
    }
    // X rows travel as ONE bulk copy per column and tile (the copy engine handles ~1 copy per 60 ns and SM: small
    // pieces made it the bottleneck), so tiles are as wide as possible: up to 512 fields = 16 consumer warps, one CTA per SM
    const int tw = std::min((nvar + 31) / 32 * 32, 512);
    const int ntile = (nvar + tw - 1) / tw;
    const size_t stagebytes = ((size_t)tw * 8 + GR_SKEW) * GR_CH + GR_REC;
    const int nstage = (int)std::min<size_t>(8, ((size_t)max_smem[dev] - 128) / stagebytes);
    if (nstage < 2) return false;
    const size_t smem = 128 + (size_t)nstage * stagebytes;
    const int64_t item_lo = row_lo / GR_R, item_hi = std::min<int64_t>(P.nitems, cdiv(row_hi, GR_R));
    if (item_lo >= item_hi) return true;
    const int64_t grid = std::max<int64_t>(1, std::min<int64_t>(item_hi - item_lo, sm_count()));
    bool fresh_split = false;
    if (P.split_lo != item_lo || P.split_hi != item_hi || P.split_grid != (int)grid) {
        fresh_split = true;
        P.split.alloc((size_t)grid + 1, s);
        gr_split_kernel<<<(unsigned)cdiv(grid + 1, 128), 128, 0, s>>>(P.wprefix.p, item_lo, item_hi, (int)grid, P.split.p);
        IB_LAUNCHED();
        P.split_lo = item_lo; P.split_hi = item_hi; P.split_grid = (int)grid;
    }
    GrArgs a{P.items.p, P.chunks.p, P.split.p, ntile, tw, nstage, dX, ldx, nvar, dY, ldy};
    // programmatic dependent launch: behind a kernel that triggers it (cons_scale_kernel) the CTAs set themselves up --
    // barriers, register hand-over, their share of the items -- while that kernel still runs; the producer waits for it
    // before the first copy of X.  Not on the call that has just launched gr_split_kernel: every thread reads its table.
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(GR_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = fresh_split ? 0 : 1;
    IB_CUDA(cudaLaunchKernelEx(&cfg, smooth_group_kernel, a));
    IB_LAUNCHED();
    return true;
}

}   // namespace ib
