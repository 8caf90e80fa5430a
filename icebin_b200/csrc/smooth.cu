// smooth.cu -- the sigma smoothing operator:  M <- smoothI * M   (RegridMatrices_Dynamic.cpp:236-248).
//
// Replaces smoothing_matrix()/Smoother (slib/icebin/smoother.cpp:8-99, smoother.hpp:13-69):
//   tuples  = unmasked ice cells of agridI that are in dimI, with centroid (x, y, elevation) and area = wM[dense id];
//   smoothI[i, j] = exp(-d2/2) * area_j / sum_j'(...)   for   d2 = sum_k ((c_j[k]-c_i[k])/sigma_k)^2 < nsigma^2 = 4.
// The RTree box query becomes a uniform bin grid with cells of 2*sigma_x by 2*sigma_y (a neighbour is always in
// the 3x3 bins around a cell); the accepted SET is identical, only the order of the normaliser's additions
// differs (upstream's order is the RTree's traversal order, i.e. implementation-defined).
//
// smoothI is never materialised (it has ~10^3 entries per row at 5 km / 50 km): one warp owns one output row,
// sweeps its candidate neighbours 32 at a time, and folds every accepted neighbour's M row into at most
// SM_CAP column slots that live in registers (lane l owns slots l, l+32, ...).  Neighbours are folded one after the
// other, so the additions of one output entry happen in a fixed order (deterministic).  Two launches: count the
// distinct columns per row, scan, then fill (columns sorted ascending inside a row, as CSR wants).
//
// Cost: ~ (candidates + accepted * nnz-per-M-row) warp steps per row; bytes are negligible (all operands sit in L2).
#include <cmath>

#include "radix_sort.cuh"
#include "scan.cuh"
#include "types.cuh"
#include "util.cuh"

namespace ib {

constexpr int SM_SPL = 8;               // slots per lane
constexpr int SM_CAP = 32 * SM_SPL;     // distinct columns one output row may have

struct Tuples {
    int32_t *d;        // dense id in dimI
    double *x, *y, *z, *area;
};

struct TupLoad {
    const int64_t *dimI_sheet;   // agridI dense -> sparse
    const int32_t *dense_of;     // dims[0] table (nullptr if identity)
    const double *em;
    __device__ uint32_t operator()(int64_t id) const
    {
        const int64_t s = dimI_sheet[id];
        if (dense_of && dense_of[s] < 0) return 0u;
        return isnan(em[s]) ? 0u : 1u;
    }
};
struct TupEmit {
    const int64_t *dimI_sheet;
    const int32_t *dense_of;
    const double *em, *centroid, *wM;
    Tuples t;
    int *err;
    __device__ void operator()(int64_t id, uint32_t prefix, uint32_t keep) const
    {
        if (!keep) return;
        const int64_t s = dimI_sheet[id];
        const int32_t d = dense_of ? dense_of[s] : (int32_t)s;
        const double a = wM[d];
        if (a == 0.) {
            if (atomicAdd(err, 1) == 0) { err[1] = (int)(s & 0xffffffff); err[2] = (int)(s >> 32); }
        }
        t.d[prefix] = d;
        t.x[prefix] = centroid[2 * id];
        t.y[prefix] = centroid[2 * id + 1];
        t.z[prefix] = em[s];
        t.area[prefix] = a;
    }
};

struct BinGeom {
    double x0, y0, bx, by;
    int nbx, nby, fx, fy;
    __host__ __device__ int binx(double x) const { return fx ? min(nbx - 1, max(0, (int)((x - x0) / bx))) : 0; }
    __host__ __device__ int biny(double y) const { return fy ? min(nby - 1, max(0, (int)((y - y0) / by))) : 0; }
};
__global__ void bin_key_kernel(Tuples t, int64_t n, BinGeom g, uint32_t *key)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) key[i] = (uint32_t)(g.biny(t.y[i]) * g.nbx + g.binx(t.x[i]));
}
__global__ void u32_to_i32_kernel(const uint32_t *a, int64_t n, int32_t *b)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) b[i] = (int32_t)a[i];
}
__global__ void row_of_tuple_kernel(const int32_t *tup_d, int64_t n, const uint32_t *cnt, uint32_t *row_cnt)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) row_cnt[tup_d[i]] = cnt[i];
}

struct SmoothArgs {
    Tuples t;
    int64_t ntup;
    BinGeom g;
    const int32_t *bin_start;     // [nbins+1]
    const uint32_t *bin_items;    // tuple indices sorted by bin (stable)
    double sx, sy, sz;
    const int32_t *rowptr, *col;  // M before smoothing (rows = dimI dense ids)
    const double *val;
    uint32_t *count;              // [ntup]   (count pass)
    const int32_t *out_rowptr;    // fill pass
    int32_t *out_col;
    double *out_val;
    int *err;
};

__device__ __forceinline__ double shfl_d(double v, int src)
{
    int lo = __shfl_sync(0xffffffffu, __double2loint(v), src);
    int hi = __shfl_sync(0xffffffffu, __double2hiint(v), src);
    return __hiloint2double(hi, lo);
}

template <bool FILL>
__global__ void __launch_bounds__(128) smooth_rows_kernel(const SmoothArgs a)
{
    const int lane = threadIdx.x & 31;
    const int64_t t0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (t0 >= a.ntup) return;
    const double x0 = a.t.x[t0], y0 = a.t.y[t0], z0 = a.t.z[t0];
    const int cx = a.g.binx(x0), cy = a.g.biny(y0);
    const int bx0 = max(0, cx - 1), bx1 = min(a.g.nbx - 1, cx + 1), by0 = max(0, cy - 1), by1 = min(a.g.nby - 1, cy + 1);

    int scol[SM_SPL];
    double sacc[SM_SPL];
#pragma unroll
    for (int s = 0; s < SM_SPL; ++s) { scol[s] = -1; sacc[s] = 0.0; }
    int nslots = 0;
    double factor = 0.0;

    // ONE sweep over the candidates: the normaliser sum_j(gauss_ij * area_j) and the un-normalised folds
    // sum_j (gauss_ij * area_j) * M[j, c] are accumulated together and the row is scaled by 1/normaliser at the end
    // (upstream scales every smoothI entry first, smoother.cpp:62-65: same value up to rounding).
    // Every lane first fetches ITS candidate's leading two matrix entries (lane-parallel, latencies overlap); the
    // serial part then only shuffles registers -- no dependent global load unless a neighbour's row has > 2 entries.
    double denom = 0.0;
    {
        const int sweep = 1;
        for (int yy = by0; yy <= by1; ++yy)
        for (int xx = bx0; xx <= bx1; ++xx) {
            const int b = yy * a.g.nbx + xx;
            const int qb = a.bin_start[b], qe = a.bin_start[b + 1];
            for (int q = qb; q < qe; q += 32) {
                const bool have = q + lane < qe;
                int tj = 0;
                double w = 0.0;
                bool acc = false;
                if (have) {
                    tj = (int)a.bin_items[q + lane];
                    const double dx = (a.t.x[tj] - x0) / a.sx, dy = (a.t.y[tj] - y0) / a.sy, dz = (a.t.z[tj] - z0) / a.sz;
                    double nd2 = 0.0;
                    nd2 += dx * dx; nd2 += dy * dy; nd2 += dz * dz;
                    if (nd2 < 4.0) { acc = true; if (FILL) w = exp(-.5 * nd2) * a.t.area[tj]; }
                }
                denom += w;           // lane-private partial; combined below in a fixed butterfly order
                int rb = 0, nent = 0, c0 = -1, c1 = -1;
                double v0 = 0.0, v1 = 0.0;
                if (acc) {
                    const int dj = a.t.d[tj];
                    rb = __ldg(a.rowptr + dj);
                    nent = __ldg(a.rowptr + dj + 1) - rb;
                    const double sij = w;                                 // gauss * area; 1/normaliser applied at the end
                    if (nent > 0) { c0 = __ldg(a.col + rb); if (FILL) v0 = sij * __ldg(a.val + rb); }
                    if (nent > 1) { c1 = __ldg(a.col + rb + 1); if (FILL) v1 = sij * __ldg(a.val + rb + 1); }
                    w = sij;
                }
                // fold one entry of a column slot table; `c`, `v` are warp-uniform
                auto fold = [&](int c, double v) {
                    bool hit = false;
#pragma unroll
                    for (int s = 0; s < SM_SPL; ++s)
                        if (scol[s] == c) { if (FILL) sacc[s] += v; hit = true; }
                    if (!__ballot_sync(0xffffffffu, hit)) {
                        if (nslots < SM_CAP && lane == (nslots & 31)) {
#pragma unroll
                            for (int s = 0; s < SM_SPL; ++s)
                                if (s == (nslots >> 5)) { scol[s] = c; sacc[s] = v; }
                        }
                        ++nslots;
                    }
                };
                // entries 0 and 1 of the batch: lanes whose neighbours feed the SAME column are combined first (their
                // values added in lane order), so the slot search runs once per distinct column of the batch instead of
                // once per neighbour -- neighbouring ice cells share their elevation classes
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int key = e == 0 ? c0 : c1;
                    const double val = e == 0 ? v0 : v1;
                    const unsigned mvalid = __ballot_sync(0xffffffffu, key >= 0);
                    if (!mvalid) continue;
                    const unsigned peers = __match_any_sync(0xffffffffu, key);
                    double gsum = 0.0;
                    if (FILL) {
                        unsigned mm = mvalid;
                        while (mm) {
                            const int k = __ffs(mm) - 1;
                            mm &= mm - 1;
                            const double vk = shfl_d(val, k);
                            if ((peers >> k) & 1u) gsum += vk;
                        }
                    }
                    unsigned leaders = __ballot_sync(0xffffffffu, key >= 0 && lane == __ffs(peers) - 1);
                    while (leaders) {
                        const int src = __ffs(leaders) - 1;
                        leaders &= leaders - 1;
                        fold(__shfl_sync(0xffffffffu, key, src), FILL ? shfl_d(gsum, src) : 0.0);
                    }
                }
                // neighbours whose M row has more than two entries: the rest, one neighbour at a time
                unsigned mlong = __ballot_sync(0xffffffffu, acc && nent > 2);
                while (mlong) {
                    const int src = __ffs(mlong) - 1;
                    mlong &= mlong - 1;
                    const int ne = __shfl_sync(0xffffffffu, nent, src);
                    const int pos0 = __shfl_sync(0xffffffffu, rb, src);
                    const double sw = FILL ? shfl_d(w, src) : 0.0;
                    for (int e = 2; e < ne; ++e) fold(__ldg(a.col + pos0 + e), FILL ? sw * __ldg(a.val + pos0 + e) : 0.0);
                }
            }
        }
        (void)sweep;
    }
    if (FILL) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) denom += shfl_d(denom, lane ^ d);
        factor = 1. / denom;
#pragma unroll
        for (int s = 0; s < SM_SPL; ++s) sacc[s] = factor * sacc[s];
    }
    if (nslots > SM_CAP) {
        if (lane == 0) atomicMax(a.err + 3, nslots);
        nslots = SM_CAP;
    }
    if (!FILL) {
        if (lane == 0) a.count[t0] = (uint32_t)nslots;
        return;
    }
    // rank the slots by column (ascending) and write the row
    int rank[SM_SPL];
#pragma unroll
    for (int s = 0; s < SM_SPL; ++s) rank[s] = 0;
    for (int u = 0; u < nslots; ++u) {
        int cu = 0;
#pragma unroll
        for (int s = 0; s < SM_SPL; ++s)
            if (s == (u >> 5)) cu = scol[s];
        cu = __shfl_sync(0xffffffffu, cu, u & 31);
#pragma unroll
        for (int s = 0; s < SM_SPL; ++s) rank[s] += (cu < scol[s]) ? 1 : 0;
    }
    const int base = a.out_rowptr[a.t.d[t0]];
#pragma unroll
    for (int s = 0; s < SM_SPL; ++s) {
        if (s * 32 + lane < nslots) {
            a.out_col[base + rank[s]] = scol[s];
            a.out_val[base + rank[s]] = sacc[s];
        }
    }
}

struct CountLoad {
    const uint32_t *c;
    __device__ uint32_t operator()(int64_t i) const { return c[i]; }
};
struct RowptrStore {
    int32_t *p;
    __device__ void operator()(int64_t i, uint32_t prefix, uint32_t) const { p[i] = (int32_t)prefix; }
};

void smooth_rows(RegridMatrices &rm, Weighted &W, Params const &p)
{
    Gcm &g = *rm.gcm;
    Sheet &sh = *g.sheets.at(rm.sheet);
    cudaStream_t s = rm.stream;
    if (sh.centroid_xy.n == 0) fail(IB_ERR_ARG, "smoothing requested but the ice grid has no centroid_xy");
    Dim &dimI = *W.dims[0];
    const int64_t nreal = (int64_t)sh.dimI_h.size();
    const int64_t nrow = W.nrow;

    // ---- tuples (smoother.cpp:69-96)
    DBuf<int> err(4, s);
    err.zero();
    ScanState<uint32_t> st;
    TupLoad tl{sh.dimI.p, dimI.identity ? nullptr : dimI.dense_of.p, rm.elevmaskI.p};
    transform_scan<uint32_t>(st, nreal, tl, AddU32{}, 0u, s);
    uint32_t ntup_u = 0;
    IB_CUDA(cudaMemcpyAsync(&ntup_u, st.total.p, sizeof ntup_u, cudaMemcpyDeviceToHost, s));
    IB_CUDA(cudaStreamSynchronize(s));
    const int64_t ntup = ntup_u;
    DBuf<int32_t> td((size_t)ntup, s);
    DBuf<double> tx((size_t)ntup, s), ty((size_t)ntup, s), tz((size_t)ntup, s), ta((size_t)ntup, s);
    Tuples T{td.p, tx.p, ty.p, tz.p, ta.p};
    if (nreal > 0)
        scan_apply<uint32_t>(st, tl, AddU32{}, 0u, TupEmit{sh.dimI.p, tl.dense_of, rm.elevmaskI.p, sh.centroid_xy.p, W.wM.p, T, err.p}, s);

    // ---- bins of nsigma*sigma in x and y over the sheet's bounding box
    BinGeom bg{};
    const double nsigma = 2.;
    bg.x0 = sh.bbox[0]; bg.y0 = sh.bbox[2];
    bg.bx = nsigma * p.sigma[0]; bg.by = nsigma * p.sigma[1];
    bg.fx = std::isfinite(bg.bx) && bg.bx > 0; bg.fy = std::isfinite(bg.by) && bg.by > 0;
    bg.nbx = bg.fx ? (int)std::min<double>(4096, std::floor((sh.bbox[1] - sh.bbox[0]) / bg.bx) + 1) : 1;
    bg.nby = bg.fy ? (int)std::min<double>(4096, std::floor((sh.bbox[3] - sh.bbox[2]) / bg.by) + 1) : 1;
    // if the box is wider than 4096 bins the last bin absorbs the rest; make sure neighbours are still found
    if (bg.fx && (sh.bbox[1] - sh.bbox[0]) / bg.bx >= 4096) { bg.bx = (sh.bbox[1] - sh.bbox[0]) / 4095; }
    if (bg.fy && (sh.bbox[3] - sh.bbox[2]) / bg.by >= 4096) { bg.by = (sh.bbox[3] - sh.bbox[2]) / 4095; }
    const int64_t nbins = (int64_t)bg.nbx * bg.nby;

    DBuf<uint32_t> k0((size_t)ntup, s), k1((size_t)ntup, s), i0((size_t)ntup, s), i1((size_t)ntup, s);
    DBuf<int32_t> ksorted((size_t)ntup, s), bin_start((size_t)nbins + 1, s);
    bin_start.zero();
    const uint32_t *items = i0.p;
    if (ntup > 0) {
        bin_key_kernel<<<(unsigned)cdiv(ntup, 256), 256, 0, s>>>(T, ntup, bg, k0.p);
        IB_LAUNCHED();
        iota_u32_kernel<<<(unsigned)cdiv(ntup, 256), 256, 0, s>>>(i0.p, ntup);
        IB_LAUNCHED();
        int w2 = radix_sort_pairs<uint32_t, uint32_t>(k0.p, i0.p, k1.p, i1.p, ntup, bits_for(nbins), s);
        items = w2 ? i1.p : i0.p;
        u32_to_i32_kernel<<<(unsigned)cdiv(ntup, 256), 256, 0, s>>>(w2 ? k1.p : k0.p, ntup, ksorted.p);
        IB_LAUNCHED();
        seg_ptr_kernel<<<(unsigned)cdiv(ntup, 256), 256, 0, s>>>(ksorted.p, ntup, nbins, bin_start.p);
        IB_LAUNCHED();
    }

    // ---- count, scan, fill
    DBuf<uint32_t> cnt((size_t)ntup, s), row_cnt((size_t)nrow, s);
    row_cnt.zero();
    SmoothArgs A{T, ntup, bg, bin_start.p, items, p.sigma[0], p.sigma[1], p.sigma[2], W.rowptr.p, W.col.p, W.val.p,
                 cnt.p, nullptr, nullptr, nullptr, err.p};
    const unsigned nblk = (unsigned)cdiv(ntup * 32, 128);
    if (ntup > 0) {
        smooth_rows_kernel<false><<<nblk, 128, 0, s>>>(A);
        IB_LAUNCHED();
        row_of_tuple_kernel<<<(unsigned)cdiv(ntup, 256), 256, 0, s>>>(T.d, ntup, cnt.p, row_cnt.p);
        IB_LAUNCHED();
    }
    DBuf<int32_t> new_rowptr((size_t)nrow + 1, s);
    new_rowptr.zero();
    int64_t new_nnz = 0;
    if (nrow > 0) {
        ScanState<uint32_t> rs;
        transform_scan<uint32_t>(rs, nrow, CountLoad{row_cnt.p}, AddU32{}, 0u, s);
        scan_apply<uint32_t>(rs, CountLoad{row_cnt.p}, AddU32{}, 0u, RowptrStore{new_rowptr.p}, s);
        uint32_t tot = 0;
        int herr[4];
        IB_CUDA(cudaMemcpyAsync(&tot, rs.total.p, sizeof tot, cudaMemcpyDeviceToHost, s));
        IB_CUDA(cudaMemcpyAsync(herr, err.p, sizeof herr, cudaMemcpyDeviceToHost, s));
        IB_CUDA(cudaStreamSynchronize(s));
        if (herr[0]) {
            long cell = (long)(((uint64_t)(uint32_t)herr[2] << 32) | (uint32_t)herr[1]);
            fail(IB_ERR_AREA, "Area of cell %ld must be non-zero\n", cell);     // smoother.cpp:88-90
        }
        if (herr[3]) fail(IB_ERR_ARG, "smoothing: an ice cell reaches %d distinct columns within 2 sigma (limit %d)", herr[3], SM_CAP);
        new_nnz = tot;
        IB_CUDA(cudaMemcpyAsync(new_rowptr.p + nrow, &tot, sizeof(int32_t), cudaMemcpyHostToDevice, s));
        IB_CUDA(cudaStreamSynchronize(s));
    }
    DBuf<int32_t> new_col((size_t)new_nnz, s);
    DBuf<double> new_val((size_t)new_nnz, s);
    if (ntup > 0 && new_nnz > 0) {
        A.out_rowptr = new_rowptr.p; A.out_col = new_col.p; A.out_val = new_val.p;
        smooth_rows_kernel<true><<<nblk, 128, 0, s>>>(A);
        IB_LAUNCHED();
    }
    IB_CUDA(cudaStreamSynchronize(s));     // the old CSR is released below
    W.rowptr = std::move(new_rowptr);
    W.col = std::move(new_col);
    W.val = std::move(new_val);
    W.nnz = new_nnz;
    W.plan = ApplyPlan();
}

}   // namespace ib
