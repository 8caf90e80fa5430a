// modele.cu -- the device step of GCMRegridder_ModelE's "mismatched" regridding.
//
// modele/GCMRegridder_ModelE.cpp:318-435 builds its matrices as products of a SMALL matrix on the GCM side (XAmvXOm:
// atmosphere cells / elevation classes <- ocean-grid cells; one entry per ocean-grid cell, since the atmosphere grid is
// the ocean grid coarsened 2:1, hntr.cpp:232-241) with a LARGE ice-side matrix of the ocean-grid regridder, whose rows
// (or columns) are first re-indexed from the ice model's ocean mask to ModelE's (crop_mvp, :285-307):
//     XAmvIp = diag(sXAm) * XAmvXOm * crop_mvp(dimXOm, dimXOp, 0, diag(1 / XOpvIp.wM) * XOpvIp)          (:367-386)
//     IpvXAm = crop_mvp(dimXOm, dimXOp, 1, diag(1 / IpvXOp.wM) * IpvXOp) * diag(XAmvXOms) * XOmvXAm     (:418-432)
// With one entry per ocean-grid cell in the small factor, both products are: relabel one side of the large matrix
// (ocean-grid cell -> its atmosphere cell, or dropped), scale every entry by a factor of its old index (and, for the
// second form, of its row), and add up the entries that now coincide.  That is K2 again: keys, stable radix sort,
// runs summed in storage order.  The small factors are computed by the host (icebin_b200/modele.py).
#include "radix_sort.cuh"
#include "scan.cuh"
#include "segsum.cuh"
#include "types.cuh"
#include "util.cuh"

namespace ib {

struct AggLoad {
    const int32_t *rowptr, *col, *map;
    int64_t nrow;
    int side;
    __device__ __forceinline__ int32_t row_of(int64_t j) const
    {
        int64_t lo = 0, hi = nrow;                  // largest r with rowptr[r] <= j
        while (hi - lo > 1) {
            const int64_t mid = (lo + hi) >> 1;
            if (rowptr[mid] <= j) lo = mid; else hi = mid;
        }
        return (int32_t)lo;
    }
    __device__ uint32_t operator()(int64_t j) const { return map[side == 0 ? row_of(j) : col[j]] >= 0 ? 1u : 0u; }
};
struct AggEmit {
    AggLoad L;
    const double *val, *factor, *other;
    int cbits;
    uint64_t *keys;
    double *vals;
    __device__ void operator()(int64_t j, uint32_t prefix, uint32_t keep) const
    {
        if (!keep) return;
        const int32_t r = L.row_of(j), c = L.col[j];
        const int32_t old = L.side == 0 ? r : c, oth = L.side == 0 ? c : r;
        const int32_t nr = L.side == 0 ? L.map[r] : r, nc = L.side == 0 ? c : L.map[c];
        double v = val[j];
        if (other) v = other[oth] * v;               // diag(1 / wM) * M  (the reference scales the large matrix first)
        if (factor) v = factor[old] * v;
        keys[prefix] = ((uint64_t)(uint32_t)nr << cbits) | (uint32_t)nc;
        vals[prefix] = v;
    }
};

// N[map[r], c] += factor[r] * other[c] * M[r, c]   (side 0)     or     N[r, map[c]] += factor[c] * other[r] * M[r, c]   (side 1)
// entries whose map is < 0 are dropped; coinciding entries are added in storage order.  Identity dims; weights zero.
std::unique_ptr<Weighted> aggregate(Weighted &W, int side, const int32_t *map_h, const double *factor_h, int64_t n_new,
                                    const double *other_h, cudaStream_t s)
{
    ensure_device();
    const int64_t n_old = side == 0 ? W.nrow : W.ncol, n_oth = side == 0 ? W.ncol : W.nrow;
    const int64_t nrow = side == 0 ? n_new : W.nrow, ncol = side == 0 ? W.ncol : n_new;
    DBuf<int32_t> map((size_t)n_old, s);
    DBuf<double> factor, other;
    map.upload(map_h, (size_t)n_old);
    if (factor_h) { factor.alloc((size_t)n_old, s); factor.upload(factor_h, (size_t)n_old); }
    if (other_h) { other.alloc((size_t)n_oth, s); other.upload(other_h, (size_t)n_oth); }
    std::unique_ptr<Weighted> N(new Weighted);
    N->stream = s; N->nrow = nrow; N->ncol = ncol;
    for (int k = 0; k < 2; ++k) {
        N->owned[k].reset(new Dim);
        N->owned[k]->identity = true;
        N->owned[k]->sparse_extent = N->owned[k]->dense_extent = k ? ncol : nrow;
        N->dims[k] = N->owned[k].get();
    }
    N->rowptr.alloc((size_t)nrow + 1, s);
    N->rowptr.zero();
    N->wM.alloc((size_t)nrow, s); N->wM.zero();
    N->Mw.alloc((size_t)ncol, s); N->Mw.zero();
    int64_t nkeep = 0, nnz = 0;
    if (W.nnz > 0) {
        AggLoad L{W.rowptr.p, W.col.p, map.p, W.nrow, side};
        ScanState<uint32_t> st;
        transform_scan<uint32_t>(st, W.nnz, L, AddU32{}, 0u, s);
        uint32_t tot = 0;
        IB_CUDA(cudaMemcpyAsync(&tot, st.total.p, sizeof tot, cudaMemcpyDeviceToHost, s));
        IB_CUDA(cudaStreamSynchronize(s));
        nkeep = tot;
        const int cbits = bits_for(ncol), rbits = bits_for(nrow);
        if (cbits + rbits > 64) fail(IB_ERR_ARG, "aggregate: matrix too large for 64-bit keys");
        DBuf<uint64_t> k0((size_t)nkeep, s), k1((size_t)nkeep, s);
        DBuf<double> v0((size_t)nkeep, s), v1((size_t)nkeep, s);
        scan_apply<uint32_t>(st, L, AddU32{}, 0u, AggEmit{L, W.val.p, factor.p, other.p, cbits, k0.p, v0.p}, s);
        const int which = radix_sort_pairs<uint64_t, double>(k0.p, v0.p, k1.p, v1.p, nkeep, cbits + rbits, s);
        const uint64_t *sk = which ? k1.p : k0.p;
        const double *sv = which ? v1.p : v0.p;
        if (nkeep > 0) {
            ScanState<uint32_t> hs;
            transform_scan<uint32_t>(hs, nkeep, HeadLoad<uint64_t>{sk}, AddU32{}, 0u, s);
            uint32_t h = 0;
            IB_CUDA(cudaMemcpyAsync(&h, hs.total.p, sizeof h, cudaMemcpyDeviceToHost, s));
            IB_CUDA(cudaStreamSynchronize(s));
            nnz = h;
            N->col.alloc((size_t)nnz, s);
            N->val.alloc((size_t)nnz, s);
            DBuf<int32_t> row((size_t)nnz, s), segstart((size_t)nnz + 1, s);
            scan_apply<uint32_t>(hs, HeadLoad<uint64_t>{sk}, AddU32{}, 0u, ReduceApply{sk, cbits, row.p, N->col.p, segstart.p}, s);
            const int32_t end = (int32_t)nkeep;
            IB_CUDA(cudaMemcpyAsync(segstart.p + nnz, &end, sizeof end, cudaMemcpyHostToDevice, s));
            seg_sum(segstart.p, nullptr, sv, nnz, N->val.p, s);
            seg_ptr_kernel<<<(unsigned)cdiv(nnz, 256), 256, 0, s>>>(row.p, nnz, nrow, N->rowptr.p);
            IB_LAUNCHED();
            IB_CUDA(cudaStreamSynchronize(s));
        }
    }
    N->nnz = nnz;
    N->conservative = false;
    N->scaled = true;
    IB_CUDA(cudaStreamSynchronize(s));
    return N;
}

}   // namespace ib
