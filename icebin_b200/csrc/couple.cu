// couple.cu -- the regridding side of the coupling step: the callers of the matrix build/apply path.
//
//   vartrans_dev       Z = O * M + b          unit/variable conversion in front of a regrid apply
//                                             (IceCoupler.cpp:203-252 gcm -> ice, :404-446 ice -> gcm)
//   e1ve0c             E1vE0c = sE1 * sum_sheets E1uX * (XvE0 - XvE1)   (e1ve0.cpp:55-106, GCMCoupler.cpp:488-515)
//   combine_chunks     per-chunk unscaled matrices -> one matrix in sparse indexing, scaled by 1/sum_chunks wM
//                                             (modele/combine_global_ec.cpp:95-262)
//
// Compiled with -fmad=false; additions run in the reference's order wherever that order is defined.
#include <algorithm>
#include <cmath>

#include "radix_sort.cuh"
#include "scan.cuh"
#include "segsum.cuh"
#include "types.cuh"
#include "util.cuh"

namespace ib {

// ---- variable transformation ---------------------------------------------------------------------------------
// Eigen evaluates dense * sparse(col-major) column by column: Z(:,k) = sum_j O(:, row_j) * val_j over the stored entries
// j of column k, ascending; then "+ b.replicate(n,1)" adds b[k].  Structural zeros are never multiplied, so a NaN in
// an input variable only reaches the outputs that name it (contracts::ALLOW_NAN variables rely on that).
__global__ void vartrans_kernel(const double *__restrict__ O, int64_t ldo, int64_t ncell, int nout, const int32_t *__restrict__ colptr,
                                const int32_t *__restrict__ row, const double *__restrict__ val, const double *__restrict__ b,
                                double *__restrict__ Z, int64_t ldz)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t i = t / nout;
    const int k = (int)(t - i * nout);
    if (i >= ncell) return;
    double acc = 0.0;
    for (int32_t j = colptr[k]; j < colptr[k + 1]; ++j) acc = acc + O[i * ldo + row[j]] * val[j];
    Z[i * ldz + k] = acc + b[k];
}

void vartrans_dev(const double *dO, int64_t ldo, int64_t ncell, int nin, int nout, const int32_t *colptr, const int32_t *row,
                  const double *val, const double *b, double *dZ, int64_t ldz, cudaStream_t s)
{
    ensure_device();
    if (nin < 0 || nout <= 0 || ncell < 0 || ldo < nin || ldz < nout) fail(IB_ERR_ARG, "vartrans: bad shape");
    const int32_t nnz = colptr[nout];
    for (int k = 0; k < nout; ++k)
        if (colptr[k] > colptr[k + 1]) fail(IB_ERR_ARG, "vartrans: colptr must be non-decreasing");
    for (int32_t j = 0; j < nnz; ++j)
        if (row[j] < 0 || row[j] >= nin) fail(IB_ERR_ARG, "vartrans: row index %d outside [0,%d)", row[j], nin);
    if (ncell == 0) return;
    DBuf<int32_t> dcp((size_t)nout + 1, s), drow((size_t)std::max(nnz, 1), s);
    DBuf<double> dval((size_t)std::max(nnz, 1), s), db((size_t)nout, s);
    dcp.upload(colptr, (size_t)nout + 1);
    drow.upload(row, (size_t)nnz);
    dval.upload(val, (size_t)nnz);
    db.upload(b, (size_t)nout);
    const int64_t tot = ncell * nout;
    vartrans_kernel<<<(unsigned)cdiv(tot, 256), 256, 0, s>>>(dO, ldo, ncell, nout, dcp.p, drow.p, dval.p, db.p, dZ, ldz);
    IB_LAUNCHED();
    IB_CUDA(cudaStreamSynchronize(s));      // the staging buffers above were filled from the caller's (pageable) arrays
}

// ---- E1vE0c ------------------------------------------------------------------------------------------------------
// Per exchange cell x the row of XuE (XvE, unscaled) holds the elevation classes that bracket the cell (<= 2 for the
// L0 parameterisation; any length is handled).  With XvE = diag(1/wM) XuE, the product E1uX * (XvE0 - XvE1) is
//   sum_x  XuE1[x,e1] * (XuE0[x,e]/wM0[x] - XuE1[x,e]/wM1[x])
// i.e. one term per (x, e1 in row1(x), e in row0(x) U row1(x)), summed over ascending x (Eigen's sparse product walks
// the right operand's columns): emit -> stable sort on (e1, e, sheet) -> sum runs in emission order -> scale by
// 1/sum_sheets Mw1[e1] -> sum over sheets (consolidate).  E indices are SPARSE ids (the sheets' dense spaces differ).
struct E1Sheet {
    const int32_t *rp1, *col1, *rp0, *col0;
    const double *val1, *val0, *wM1, *wM0;
    const int64_t *sp1, *sp0;      // dense -> sparse E ids (nullptr: identity)
};
__device__ __forceinline__ int64_t e1_sparse(const int64_t *sp, int32_t d) { return sp ? sp[d] : (int64_t)d; }

struct E1CountLoad {
    E1Sheet h;
    __device__ uint32_t operator()(int64_t x) const
    {
        const int32_t b1 = h.rp1[x], e1 = h.rp1[x + 1], b0 = h.rp0[x], e0 = h.rp0[x + 1];
        if (e1 == b1) return 0u;
        uint32_t nun = (uint32_t)(e0 - b0);
        for (int32_t j = b1; j < e1; ++j) {
            const int64_t s = e1_sparse(h.sp1, h.col1[j]);
            bool found = false;
            for (int32_t q = b0; q < e0; ++q) found = found || (e1_sparse(h.sp0, h.col0[q]) == s);
            nun += found ? 0u : 1u;
        }
        return (uint32_t)(e1 - b1) * nun;
    }
};
struct E1Emit {
    E1Sheet h;
    uint64_t nE, nsheet, sheet;
    int64_t base;
    uint64_t *key;
    double *val;
    __device__ void operator()(int64_t x, uint32_t prefix, uint32_t cnt) const
    {
        if (!cnt) return;
        const int32_t b1 = h.rp1[x], e1 = h.rp1[x + 1], b0 = h.rp0[x], e0 = h.rp0[x + 1];
        const double s0 = 1. / h.wM0[x], s1 = 1. / h.wM1[x];      // sXuE0, sXuE1 (e1ve0.cpp:78-79)
        int64_t p = base + prefix;
        for (int32_t j = b1; j < e1; ++j) {                       // e1 of E1uX(:, x)
            const uint64_t er = (uint64_t)e1_sparse(h.sp1, h.col1[j]);
            const double u = h.val1[j];
            // columns of (XvE0 - XvE1)(x, :): entries of row0 (minus the matching entry of row1), then row1-only entries
            for (int32_t q = b0; q < e0; ++q) {
                const int64_t ec = e1_sparse(h.sp0, h.col0[q]);
                double d = s0 * h.val0[q];
                for (int32_t r = b1; r < e1; ++r)
                    if (e1_sparse(h.sp1, h.col1[r]) == ec) d = d - s1 * h.val1[r];
                key[p] = ((er * nE + (uint64_t)ec) * nsheet) + sheet;
                val[p] = u * d;
                ++p;
            }
            for (int32_t r = b1; r < e1; ++r) {
                const int64_t ec = e1_sparse(h.sp1, h.col1[r]);
                bool found = false;
                for (int32_t q = b0; q < e0; ++q) found = found || (e1_sparse(h.sp0, h.col0[q]) == ec);
                if (found) continue;
                key[p] = ((er * nE + (uint64_t)ec) * nsheet) + sheet;
                val[p] = u * (0. - s1 * h.val1[r]);
                ++p;
            }
        }
    }
};
__global__ void e1_scatter_add_kernel(const double *Mw, const int64_t *sp, int64_t n, double *sE1)
{
    const int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (d < n) sE1[e1_sparse(sp, (int32_t)d)] += Mw[d];      // one dense id per sparse id inside a sheet: no conflicts
}
// runs of equal (e1, e, sheet) -> value * (1/sE1[e1]) keyed by (e1, e)
__global__ void e1_scale_kernel(const uint64_t *key3, const int32_t *segstart, const double *sums, int64_t n, uint64_t nE, uint64_t nsheet,
                                const double *sE1, uint64_t *key2, double *out)
{
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    const uint64_t k = key3[segstart[g]] / nsheet;
    const double sc = 1. / sE1[k / nE];                       // "sE1(i) = 1. / sE1(i)" then "value *= sE1(iE1)"
    key2[g] = k;
    out[g] = sums[g] * sc;
}
__global__ void e1_heads_kernel(const uint64_t *key, const int32_t *segstart, int64_t n, uint64_t nE, int64_t *rows, int64_t *cols)
{
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    const uint64_t k = key[segstart[g]];
    rows[g] = (int64_t)(k / nE);
    cols[g] = (int64_t)(k % nE);
}
struct SegStartApply {
    int32_t *segstart;
    __device__ void operator()(int64_t i, uint32_t prefix, uint32_t head) const { if (head) segstart[prefix] = (int32_t)i; }
};

// sorted keys -> (number of runs, start of every run + n)
static int64_t run_starts(const uint64_t *sk, int64_t n, DBuf<int32_t> &segstart, cudaStream_t s)
{
    ScanState<uint32_t> hs;
    transform_scan<uint32_t>(hs, n, HeadLoad<uint64_t>{sk}, AddU32{}, 0u, s);
    uint32_t h = 0;
    IB_CUDA(cudaMemcpyAsync(&h, hs.total.p, sizeof h, cudaMemcpyDeviceToHost, s));
    IB_CUDA(cudaStreamSynchronize(s));
    segstart.alloc((size_t)h + 1, s);
    scan_apply<uint32_t>(hs, HeadLoad<uint64_t>{sk}, AddU32{}, 0u, SegStartApply{segstart.p}, s);
    const int32_t n32 = (int32_t)n;
    IB_CUDA(cudaMemcpyAsync(segstart.p + h, &n32, sizeof n32, cudaMemcpyHostToDevice, s));
    IB_CUDA(cudaStreamSynchronize(s));
    return h;
}

void e1ve0c(int nsheet, Weighted *const *XuE1, Weighted *const *XuE0, int64_t nE, std::vector<int64_t> &rows, std::vector<int64_t> &cols,
            std::vector<double> &vals, cudaStream_t s)
{
    ensure_device();
    rows.clear(); cols.clear(); vals.clear();
    if (nsheet <= 0 || nE <= 0) fail(IB_ERR_ARG, "e1ve0c: bad argument");
    if ((double)nE * (double)nE * (double)nsheet >= 18446744073709551615.0) fail(IB_ERR_ARG, "e1ve0c: key space too large");
    std::vector<E1Sheet> hs((size_t)nsheet);
    std::vector<ScanState<uint32_t>> st((size_t)nsheet);
    std::vector<int64_t> base((size_t)nsheet + 1, 0);
    DBuf<double> sE1((size_t)nE, s);
    sE1.zero();
    for (int i = 0; i < nsheet; ++i) {
        Weighted &W1 = *XuE1[i], &W0 = *XuE0[i];
        if (W1.nrow != W0.nrow) fail(IB_ERR_DIM, "e1ve0c: sheet %d: XuE1 has %ld exchange cells, XuE0 %ld", i, (long)W1.nrow, (long)W0.nrow);
        for (Weighted *W : {&W1, &W0})
            if (!W->dims[1] || W->dims[1]->sparse_extent > nE) fail(IB_ERR_DIM, "e1ve0c: sheet %d: E dimension does not fit nE = %ld", i, (long)nE);
        E1Sheet &h = hs[i];
        h.rp1 = W1.rowptr.p; h.col1 = W1.col.p; h.val1 = W1.val.p; h.wM1 = W1.wM.p;
        h.rp0 = W0.rowptr.p; h.col0 = W0.col.p; h.val0 = W0.val.p; h.wM0 = W0.wM.p;
        h.sp1 = W1.dims[1]->identity ? nullptr : W1.dims[1]->sparse_of.p;
        h.sp0 = W0.dims[1]->identity ? nullptr : W0.dims[1]->sparse_of.p;
        uint32_t tot = 0;
        if (W1.nrow > 0) {
            transform_scan<uint32_t>(st[i], W1.nrow, E1CountLoad{h}, AddU32{}, 0u, s);
            IB_CUDA(cudaMemcpyAsync(&tot, st[i].total.p, sizeof tot, cudaMemcpyDeviceToHost, s));
            IB_CUDA(cudaStreamSynchronize(s));
        }
        base[i + 1] = base[i] + tot;
        if (W1.ncol > 0) {      // sE1 += Mw of the sparsified XuE1, sheet after sheet (e1ve0.cpp:93)
            e1_scatter_add_kernel<<<(unsigned)cdiv(W1.ncol, 256), 256, 0, s>>>(W1.Mw.p, h.sp1, W1.ncol, sE1.p);
            IB_LAUNCHED();
        }
    }
    const int64_t nt = base[nsheet];
    if (nt == 0) return;
    if (nt >= (int64_t(1) << 31)) fail(IB_ERR_ARG, "e1ve0c: too many terms");
    DBuf<uint64_t> k0((size_t)nt, s), k1((size_t)nt, s);
    DBuf<double> v0((size_t)nt, s), v1((size_t)nt, s);
    for (int i = 0; i < nsheet; ++i)
        if (base[i + 1] > base[i])
            scan_apply<uint32_t>(st[i], E1CountLoad{hs[i]}, AddU32{}, 0u,
                                 E1Emit{hs[i], (uint64_t)nE, (uint64_t)nsheet, (uint64_t)i, base[i], k0.p, v0.p}, s);
    // level 1: per sheet sums of equal (e1, e)  [the Eigen product of that sheet], scaled by 1/sE1[e1]
    int kb = 1;
    while (kb < 64 && (double)(uint64_t(1) << kb) < (double)nE * (double)nE * (double)nsheet) ++kb;
    int w = radix_sort_pairs<uint64_t, double>(k0.p, v0.p, k1.p, v1.p, nt, kb, s);
    const uint64_t *sk = w ? k1.p : k0.p;
    const double *sv = w ? v1.p : v0.p;
    DBuf<int32_t> seg1;
    const int64_t n1 = run_starts(sk, nt, seg1, s);
    DBuf<double> sum1((size_t)n1, s), sc1((size_t)n1, s);
    DBuf<uint64_t> key2((size_t)n1, s);
    seg_sum(seg1.p, nullptr, sv, n1, sum1.p, s);
    e1_scale_kernel<<<(unsigned)cdiv(n1, 256), 256, 0, s>>>(sk, seg1.p, sum1.p, n1, (uint64_t)nE, (uint64_t)nsheet, sE1.p, key2.p, sc1.p);
    IB_LAUNCHED();
    // level 2: consolidate over sheets (keys are already sorted; equal keys are adjacent, in sheet order)
    DBuf<int32_t> seg2;
    const int64_t n2 = run_starts(key2.p, n1, seg2, s);
    DBuf<double> out((size_t)n2, s);
    DBuf<int64_t> drow((size_t)n2, s), dcol((size_t)n2, s);
    seg_sum(seg2.p, nullptr, sc1.p, n2, out.p, s);
    e1_heads_kernel<<<(unsigned)cdiv(n2, 256), 256, 0, s>>>(key2.p, seg2.p, n2, (uint64_t)nE, drow.p, dcol.p);
    IB_LAUNCHED();
    rows.resize((size_t)n2); cols.resize((size_t)n2); vals.resize((size_t)n2);
    drow.download(rows.data(), (size_t)n2);
    dcol.download(cols.data(), (size_t)n2);
    out.download(vals.data(), (size_t)n2);
    IB_CUDA(cudaStreamSynchronize(s));
}

// ---- combine_chunks (modele/combine_global_ec.cpp:95-262) ----------------------------------------------------------
// The reference computes global matrices in chunks (ranges of GCM cells, modele/global_ec.cpp:818-893; here also one chunk
// per ice sheet or per GPU domain) and concatenates them: every chunk's entries are relabelled with SPARSE indices,
// weights wM are accumulated per sparse row over all chunks first (sM_s = 1 / sum wM), and with `scale` every entry is
// multiplied by sM_s of its row.  Entries and weights are concatenated in chunk order, not merged (:233 checks the count).
__global__ void expand_rows_kernel(const int32_t *rowptr, int64_t nrow, int32_t *row);      // apply.cu
__global__ void cc_entries_kernel(const int32_t *row, const int32_t *col, const double *val, int64_t nnz, const int64_t *sp0,
                                  const int64_t *sp1, const double *sM, int scale, int64_t *orow, int64_t *ocol, double *oval)
{
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nnz) return;
    const int64_t rs = e1_sparse(sp0, row[j]);
    orow[j] = rs;
    ocol[j] = e1_sparse(sp1, col[j]);
    oval[j] = scale ? val[j] * (1. / sM[rs]) : val[j];
}
__global__ void cc_ids_kernel(const int64_t *sp, int64_t n, int64_t *out)
{
    const int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (d < n) out[d] = e1_sparse(sp, (int32_t)d);
}

void combine_chunks(int nchunk, Weighted *const *W, bool scale, int64_t *M_row, int64_t *M_col, double *M_val, int64_t *wM_index,
                    double *wM_val, int64_t *Mw_index, double *Mw_val, cudaStream_t s)
{
    ensure_device();
    if (nchunk <= 0) fail(IB_ERR_ARG, "combine_chunks: no chunks");
    for (int i = 0; i < nchunk; ++i) {
        if (!W[i]->dims[0] || !W[i]->dims[1]) fail(IB_ERR_ARG, "combine_chunks: chunk %d has no dims", i);
        if (W[i]->scaled) fail(IB_ERR_ARG, "combine_chunks: chunk %d is scaled; chunks must be built with scale = false (global_ec.cpp:273)", i);
        for (int k = 0; k < 2; ++k)
            if (W[i]->dims[k]->sparse_extent != W[0]->dims[k]->sparse_extent)
                fail(IB_ERR_DIM, "combine_chunks: chunk %d lives in a different sparse space", i);
    }
    const int64_t n0 = W[0]->dims[0]->sparse_extent;
    if (n0 <= 0) fail(IB_ERR_ARG, "combine_chunks: unknown sparse extent");
    DBuf<double> sM((size_t)n0, s);
    sM.zero();
    for (int i = 0; i < nchunk; ++i) {       // weights first, chunk after chunk (combine_global_ec.cpp:157-174)
        Weighted &w = *W[i];
        if (w.nrow > 0) {
            e1_scatter_add_kernel<<<(unsigned)cdiv(w.nrow, 256), 256, 0, s>>>(w.wM.p, w.dims[0]->identity ? nullptr : w.dims[0]->sparse_of.p, w.nrow, sM.p);
            IB_LAUNCHED();
        }
    }
    int64_t o_nnz = 0, o_row = 0, o_col = 0;
    for (int i = 0; i < nchunk; ++i) {
        Weighted &w = *W[i];
        const int64_t *sp0 = w.dims[0]->identity ? nullptr : w.dims[0]->sparse_of.p;
        const int64_t *sp1 = w.dims[1]->identity ? nullptr : w.dims[1]->sparse_of.p;
        if (w.nnz > 0) {
            DBuf<int32_t> row((size_t)w.nnz, s);
            DBuf<int64_t> dr((size_t)w.nnz, s), dc((size_t)w.nnz, s);
            DBuf<double> dv((size_t)w.nnz, s);
            expand_rows_kernel<<<(unsigned)cdiv(w.nrow, 128), 128, 0, s>>>(w.rowptr.p, w.nrow, row.p);
            IB_LAUNCHED();
            cc_entries_kernel<<<(unsigned)cdiv(w.nnz, 256), 256, 0, s>>>(row.p, w.col.p, w.val.p, w.nnz, sp0, sp1, sM.p, scale ? 1 : 0, dr.p, dc.p, dv.p);
            IB_LAUNCHED();
            dr.download(M_row + o_nnz, (size_t)w.nnz);
            dc.download(M_col + o_nnz, (size_t)w.nnz);
            dv.download(M_val + o_nnz, (size_t)w.nnz);
            IB_CUDA(cudaStreamSynchronize(s));
        }
        for (int k = 0; k < 2; ++k) {
            const int64_t n = k == 0 ? w.nrow : w.ncol;
            if (n == 0) continue;
            DBuf<int64_t> ids((size_t)n, s);
            cc_ids_kernel<<<(unsigned)cdiv(n, 256), 256, 0, s>>>(k == 0 ? sp0 : sp1, n, ids.p);
            IB_LAUNCHED();
            ids.download(k == 0 ? wM_index + o_row : Mw_index + o_col, (size_t)n);
            (k == 0 ? w.wM : w.Mw).download(k == 0 ? wM_val + o_row : Mw_val + o_col, (size_t)n);
            IB_CUDA(cudaStreamSynchronize(s));
        }
        o_nnz += w.nnz; o_row += w.nrow; o_col += w.ncol;
    }
}

}   // namespace ib
