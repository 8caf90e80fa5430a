// apply_group.cu -- Y = M X for TALL matrices whose neighbouring rows name mostly the same columns: the sigma-smoothed
// IvE / IvA (RegridMatrices_Dynamic.cpp:236-248, smoother.cpp:43-99): ~16 nonzeros per row, and the ~26 distinct columns
// of 16 consecutive ice cells are the elevation classes their common neighbourhood feeds.
//
// Why not the other kernels.  X is small (E or A cells: L2-resident), Y is large (ice cells).  A kernel that reads one X
// row per nonzero moves 16x the output through L2 (the chip's L2 delivers ~12 TB/s: 4.5 GB -> 0.4 ms, what
// spmm_xs_kernel measured); keeping X rows in shared memory still costs one 16-byte shared-memory read per nonzero and
// field pair (36 TB/s of LDS bandwidth: 125 us).  So the X values have to be reused from REGISTERS:
//   item  = 16 consecutive rows; its distinct columns, ascending, cut into chunks of <= 8 columns;
//   chunk = the X rows of its columns (ONE 1-D bulk async copy each: 4 KB at 512 fields) + a dense 8 x 16 coefficient
//           table with a 16-bit presence mask per column, prepared once per matrix; a chunk is one stage of a
//           shared-memory ring fed by a producer warp;
//   a consumer warp owns a 64-field strip of the item: 16 running sums of one double2 per lane; for every column ONE
//   shared-memory read of x serves the (up to) 16 rows that name it.  Rows that do not name a column carry a zero in the
//   table, which adds nothing while x is finite; where a lane holds a NaN / Inf the presence bits take over, so a NaN in
//   X reaches only the rows that name its cell (structural zeros are never multiplied into a sum).
// Each row's terms are added in ascending column order, with fused multiply-add: fp64 FMA rate is what bounds this
// kernel next to L2 (1.1 M nonzeros x 512 fields = 0.56 G multiply-adds; B200 measured 18.5 T DFMA/s, half of that for
// separate DMUL + DADD), so results agree with the oracle to rounding (1e-12 gate), not bit for bit.
// (First version: every warp its own pipeline with 512-byte copies -- 0.9 M copies per apply, one per ~60 ns and SM: the
// copy engine, not L2 or the fp64 pipe, set the pace at 380 us.)
//
// Roofline: HBM (Y written once: 12*nnz + 8*nvar*(nrow + ncol) algorithmic bytes); next bounds: L2 -> SM traffic of the
// X rows (~1.6x the output at 5 km) and the fp64 pipe.
#include <algorithm>
#include <mutex>

#include "scan.cuh"
#include "tma.cuh"
#include "types.cuh"

namespace ib {

constexpr int GR_R = 16;                 // rows per item
constexpr int GR_CH = 8;                 // columns per chunk
constexpr int GR_MAXD = 128;             // distinct columns per item the plan accepts
// chunk record: ncols, pad[3] | cols[CH] | mask[CH] (uint16) + pad | coef[CH][16]
constexpr int GR_OFF_COLS = 16, GR_OFF_MASK = GR_OFF_COLS + 4 * GR_CH, GR_OFF_COEF = GR_OFF_MASK + 2 * GR_CH + 16;
constexpr int GR_REC = GR_OFF_COEF + 8 * GR_CH * GR_R;
static_assert(GR_OFF_COEF % 16 == 0 && GR_REC % 16 == 0, "bulk copies move multiples of 16 bytes");

// ---- plan: one thread per item, twice (count -> scan -> emit)
__device__ int gr_distinct(const int32_t *rowptr, const int32_t *col, int64_t r0, int64_t r1, int32_t *list)
{
    int n = 0;
    for (int32_t j = rowptr[r0]; j < rowptr[r1]; ++j) {
        const int32_t c = col[j];
        int lo = 0, hi = n;                      // sorted insert
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (list[mid] < c) lo = mid + 1; else hi = mid; }
        if (lo < n && list[lo] == c) continue;
        if (n == GR_MAXD) return GR_MAXD + 1;
        for (int k = n; k > lo; --k) list[k] = list[k - 1];
        list[lo] = c;
        ++n;
    }
    return n;
}
struct GrCountLoad {
    const int32_t *rowptr, *col;
    int64_t nrow;
    int *too_wide;
    __device__ uint32_t operator()(int64_t it) const
    {
        int32_t list[GR_MAXD];
        const int64_t r0 = it * GR_R, r1 = min(nrow, r0 + GR_R);
        const int n = gr_distinct(rowptr, col, r0, r1, list);
        if (n > GR_MAXD) { atomicExch(too_wide, 1); return 1u; }
        return (uint32_t)max(1, (n + GR_CH - 1) / GR_CH);
    }
};
struct GrEmit {
    const int32_t *rowptr, *col;
    const double *val;
    int64_t nrow;
    int4 *items;
    unsigned char *chunks;
    __device__ void operator()(int64_t it, uint32_t prefix, uint32_t nch) const
    {
        int32_t list[GR_MAXD];
        const int64_t r0 = it * GR_R, r1 = min(nrow, r0 + GR_R);
        const int n = gr_distinct(rowptr, col, r0, r1, list);
        items[it] = make_int4((int)r0, (int)prefix, (int)(prefix + nch), (int)(r1 - r0));
        for (uint32_t q = 0; q < nch; ++q) {
            unsigned char *rec = chunks + (size_t)(prefix + q) * GR_REC;
            const int c0 = (int)q * GR_CH, nc = max(0, min(GR_CH, n - c0));
            reinterpret_cast<int32_t *>(rec)[0] = nc;
            int32_t *cols = reinterpret_cast<int32_t *>(rec + GR_OFF_COLS);
            uint16_t *mask = reinterpret_cast<uint16_t *>(rec + GR_OFF_MASK);
            double *coef = reinterpret_cast<double *>(rec + GR_OFF_COEF);
            for (int k = 0; k < GR_CH; ++k) { cols[k] = k < nc ? list[c0 + k] : 0; mask[k] = 0; }
            for (int k = 0; k < GR_CH * GR_R; ++k) coef[k] = 0.0;
        }
        for (int64_t r = r0; r < r1; ++r)
            for (int32_t j = rowptr[r]; j < rowptr[r + 1]; ++j) {
                const int32_t c = col[j];
                int lo = 0, hi = n;
                while (lo < hi) { const int mid = (lo + hi) >> 1; if (list[mid] < c) lo = mid + 1; else hi = mid; }
                unsigned char *rec = chunks + (size_t)(prefix + lo / GR_CH) * GR_REC;
                const int k = lo % GR_CH, rr = (int)(r - r0);
                // duplicate (row, column) entries of a caller-supplied CSR add up, in storage order
                double *cf = reinterpret_cast<double *>(rec + GR_OFF_COEF) + k * GR_R + rr;
                uint16_t *mk = reinterpret_cast<uint16_t *>(rec + GR_OFF_MASK) + k;
                *cf = (*mk >> rr & 1) ? *cf + val[j] : val[j];
                *mk |= (uint16_t)(1u << rr);
            }
    }
};

static bool make_group_plan(Weighted &W, cudaStream_t s)
{
    GroupPlan &P = W.group_plan;
    if (P.tried) return P.ok;
    P.tried = true;
    if (W.nrow < 4 * GR_R || W.nnz < 4 * W.nrow || W.nrow < 4 * W.ncol || W.nnz >= (int64_t(1) << 31)) return false;
    const int64_t nitems = cdiv(W.nrow, GR_R);
    DBuf<int> wide(1, s);
    wide.zero();
    ScanState<uint32_t> st;
    GrCountLoad L{W.rowptr.p, W.col.p, W.nrow, wide.p};
    transform_scan<uint32_t>(st, nitems, L, AddU32{}, 0u, s);
    uint32_t nch = 0;
    int hwide = 0;
    IB_CUDA(cudaMemcpyAsync(&nch, st.total.p, sizeof nch, cudaMemcpyDeviceToHost, s));
    wide.download(&hwide, 1);
    IB_CUDA(cudaStreamSynchronize(s));
    // too many distinct columns per 16 rows, or tables much larger than the matrix: the rows do not share columns
    if (hwide || (int64_t)nch * GR_REC > 6 * 12 * W.nnz + (64 << 10)) return false;
    P.nitems = nitems;
    P.nchunks = nch;
    P.items.alloc((size_t)nitems, s);
    P.chunks.alloc((size_t)nch * GR_REC, s);
    scan_apply<uint32_t>(st, L, AddU32{}, 0u, GrEmit{W.rowptr.p, W.col.p, W.val.p, W.nrow, P.items.p, P.chunks.p}, s);
    IB_CUDA(cudaStreamSynchronize(s));
    P.ok = true;
    return true;
}

// ---- kernel: the ring of spmm_stream_kernel -- a producer warp feeds stages (= chunks) to up to eight consumer warps
struct GrArgs {
    const int4 *items;
    const unsigned char *chunks;
    int64_t nitems;
    int ntile, tw, nstage;
    const double *X;
    int64_t ldx;
    int nvar;
    double *Y;
    int64_t ldy;
    int row_lo, row_hi;
};

__global__ void __launch_bounds__(288, 2) smooth_group_kernel(const GrArgs a)
{
    extern __shared__ __align__(128) unsigned char gr_raw[];
    uint64_t *full = reinterpret_cast<uint64_t *>(gr_raw);
    uint64_t *empty = full + a.nstage;
    unsigned char *data = gr_raw + 128;
    const uint32_t rowbytes = (uint32_t)a.tw * 8u;
    const uint32_t stagebytes = rowbytes * GR_CH + GR_REC;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ncw = (int)(blockDim.x >> 5) - 1;
    if (threadIdx.x == 0) {
        for (int s = 0; s < a.nstage; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, ncw); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // contiguous share of the (item, tile) units, item-major: a CTA walks neighbouring items (their X rows overlap in L2)
    const int64_t nunits = a.nitems * a.ntile;
    const int64_t u0 = nunits * blockIdx.x / gridDim.x, u1 = nunits * (blockIdx.x + 1) / gridDim.x;
    int stage = 0;
    uint32_t phase = 0;
    auto in_block = [&](const int4 &it) { return it.x >= a.row_lo && it.x < a.row_hi; };
    if (warp == ncw) {
        // ===== producer: one stage per chunk; lane k < ncols copies the X row of column k, lane 31 the chunk record.
        // The column ids of a stage are loaded one stage ahead, so the only thing the producer ever waits for is a
        // free stage.
        int64_t u = u0;
        int4 it = make_int4(0, 0, 0, 0);
        while (u < u1 && !in_block(it = __ldg(a.items + u / a.ntile))) ++u;
        if (u >= u1) return;
        int q = it.y;
        auto meta = [&](int qq, int &nc, int32_t &c) {
            const unsigned char *rec = a.chunks + (size_t)qq * GR_REC;
            nc = __ldg(reinterpret_cast<const int32_t *>(rec));
            c = lane < GR_CH ? __ldg(reinterpret_cast<const int32_t *>(rec + GR_OFF_COLS) + lane) : 0;
        };
        int nc;
        int32_t c;
        meta(q, nc, c);
        while (u < u1) {
            // the stage after this one
            int64_t un = u;
            int4 itn = it;
            int qn = q + 1;
            if (qn >= it.z) {
                ++un;
                while (un < u1 && !in_block(itn = __ldg(a.items + un / a.ntile))) ++un;
                qn = itn.y;
            }
            int ncn = 0;
            int32_t cn = 0;
            if (un < u1) meta(qn, ncn, cn);
            const int f0 = (int)(u % a.ntile) * a.tw;
            const uint32_t tb = (uint32_t)min(a.tw, a.nvar - f0) * 8u;
            unsigned char *sp = data + (size_t)stage * stagebytes;
            if (lane == 0) {
                mbar_wait(empty + stage, phase ^ 1u);
                mbar_expect_tx(full + stage, tb * (uint32_t)nc + GR_REC);
            }
            __syncwarp();
            if (lane < nc) bulk_g2s(sp + (size_t)lane * rowbytes, a.X + (int64_t)c * a.ldx + f0, tb, full + stage);
            if (lane == 31) bulk_g2s(sp + (size_t)GR_CH * rowbytes, a.chunks + (size_t)q * GR_REC, GR_REC, full + stage);
            if (++stage == a.nstage) { stage = 0; phase ^= 1u; }
            u = un; it = itn; q = qn; nc = ncn; c = cn;
        }
        return;
    }
    // ===== consumers: lane owns fields f0 + 64*warp + 2*lane, +1 of all 16 rows of the item
    const int fo = 64 * warp + 2 * lane;
    double2 acc[GR_R];
#pragma unroll
    for (int r = 0; r < GR_R; ++r) acc[r] = make_double2(0.0, 0.0);
    int4 it_next = u0 < u1 ? __ldg(a.items + u0 / a.ntile) : make_int4(0, 0, 0, 0);
    for (int64_t u = u0; u < u1; ++u) {
        const int4 it = it_next;
        if (u + 1 < u1) it_next = __ldg(a.items + (u + 1) / a.ntile);
        if (!in_block(it)) continue;
        const int f0 = (int)(u % a.ntile) * a.tw;
        for (int q = it.y; q < it.z; ++q) {
            mbar_wait(full + stage, phase);
            const unsigned char *sb = data + (size_t)stage * stagebytes;
            const unsigned char *rec = sb + (size_t)GR_CH * rowbytes;
#ifdef IB_GROUP_SKIP_MATH       // tools/: delivery rate of the ring alone (results are wrong)
            const int nc = 0;
#else
            const int nc = *reinterpret_cast<const int32_t *>(rec);
#endif
            const unsigned char *xs = sb + (size_t)fo * 8;
#pragma unroll 2
            for (int c = 0; c < nc; ++c) {
                const double2 x = *reinterpret_cast<const double2 *>(xs + (size_t)c * rowbytes);
                const double2 *cf = reinterpret_cast<const double2 *>(rec + GR_OFF_COEF + c * (8 * GR_R));
                // A row that does not name this column has coefficient 0 in the table; 0 * x adds nothing as long as x is
                // finite.  Only if some lane holds a NaN / Inf here (a masked field value) do the presence bits decide.
                const bool finite = ((__double2hiint(x.x) & 0x7ff00000) != 0x7ff00000) && ((__double2hiint(x.y) & 0x7ff00000) != 0x7ff00000);
                if (__all_sync(0xffffffffu, finite)) {
#pragma unroll
                    for (int r2 = 0; r2 < GR_R / 2; ++r2) {
                        const double2 v = cf[r2];                         // coefficients of rows 2*r2, 2*r2 + 1
                        acc[2 * r2].x = fma(v.x, x.x, acc[2 * r2].x); acc[2 * r2].y = fma(v.x, x.y, acc[2 * r2].y);
                        acc[2 * r2 + 1].x = fma(v.y, x.x, acc[2 * r2 + 1].x); acc[2 * r2 + 1].y = fma(v.y, x.y, acc[2 * r2 + 1].y);
                    }
                } else {
                    const uint32_t m = *reinterpret_cast<const uint16_t *>(rec + GR_OFF_MASK + 2 * c);
#pragma unroll 1
                    for (int r2 = 0; r2 < GR_R / 2; ++r2) {
                        const double2 v = cf[r2];
                        const bool p0 = m >> (2 * r2) & 1u, p1 = m >> (2 * r2 + 1) & 1u;
#pragma unroll
                        for (int k = 0; k < GR_R / 2; ++k) {              // static register indices: select the pair k == r2
                            if (k == r2) {
                                if (p0) { acc[2 * k].x = fma(v.x, x.x, acc[2 * k].x); acc[2 * k].y = fma(v.x, x.y, acc[2 * k].y); }
                                if (p1) { acc[2 * k + 1].x = fma(v.y, x.x, acc[2 * k + 1].x); acc[2 * k + 1].y = fma(v.y, x.y, acc[2 * k + 1].y); }
                            }
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + stage);
            if (++stage == a.nstage) { stage = 0; phase ^= 1u; }
        }
        const int f = f0 + fo;
#ifdef IB_GROUP_SKIP_STORE      // tools/: the ring without its output (results are wrong)
        if (acc[0].x == 12345.678 && fo < a.tw && f < a.nvar) {
#else
        if (fo < a.tw && f < a.nvar) {
#endif
            double *y = a.Y + (int64_t)it.x * a.ldy + f;
#pragma unroll
            for (int r = 0; r < GR_R; ++r)
                if (r < it.w) *reinterpret_cast<double2 *>(y + (int64_t)r * a.ldy) = acc[r];
        }
#pragma unroll
        for (int r = 0; r < GR_R; ++r) acc[r] = make_double2(0.0, 0.0);
    }
}

// Applies W through the group kernel; false (nothing launched) if the matrix does not fit it.
bool apply_group(Weighted &W, const double *dX, int64_t ldx, int nvar, double *dY, int64_t ldy, int row_lo, int row_hi, cudaStream_t s)
{
    if (nvar < 32 || nvar % 2 || ldx % 2 || ldy % 2 || (uintptr_t)dX % 16 || (uintptr_t)dY % 16) return false;
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
    // X rows travel as ONE bulk copy per column and tile (the copy engine handles ~1 copy per 50 ns and SM: 512-byte
    // pieces made it the bottleneck), so tiles are as wide as possible: up to 512 fields = 8 consumer warps
    const int tw = std::min((nvar + 63) / 64 * 64, 512);
    const int ntile = (nvar + tw - 1) / tw;
    const int ncw = tw / 64;
    const size_t stagebytes = (size_t)tw * 8 * GR_CH + GR_REC;
    // two CTAs per SM: 16 consumer warps keep the fp64 pipe busy (with 8, a warp's own DFMA issue interval left it half idle)
    const int nstage = (int)std::min<size_t>(8, ((size_t)(max_smem[dev] + 1024) / 2 - 1024 - 128) / stagebytes);
    if (nstage < 2) return false;
    const size_t smem = 128 + (size_t)nstage * stagebytes;
    GrArgs a{P.items.p, P.chunks.p, P.nitems, ntile, tw, nstage, dX, ldx, nvar, dY, ldy, row_lo, row_hi};
    int occ = 0;
    IB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, smooth_group_kernel, 32 * (ncw + 1), smem));
    if (occ < 1) return false;
    const int64_t grid = std::max<int64_t>(1, std::min<int64_t>(P.nitems * ntile, (int64_t)sm_count() * std::min(occ, 2)));
    smooth_group_kernel<<<(unsigned)grid, 32 * (ncw + 1), smem, s>>>(a);
    IB_LAUNCHED();
    return true;
}

}   // namespace ib
