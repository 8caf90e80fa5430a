// apply.cu -- batched fp64 CSR SpMM  Y[nrow x nvar] = M * X[ncol x nvar]  with fields stored cell-major
// (the nvar values of one cell are contiguous), so every nonzero drives one coalesced vector read of X.
//
// Replaces ibmisc linear::Weighted_Eigen::apply()/apply_M() (call sites merge_topo.cpp:65, modele/icebin22m.cpp:153)
// and the explicit Eigen products of the coupling step (IceCoupler.cpp:237-238, 445-446).
//
// Compiled with -fmad=false: a row is accumulated as acc = acc + v*x over ascending columns, which is exactly the
// order and rounding of Eigen's column-major sparse * dense on the CPU, so rows that are not split agree with the
// oracle bit for bit.  Rows longer than APPLY_CHUNK nonzeros are cut into chunks whose partial sums are added in
// chunk order by a second kernel (deterministic; differs from the CPU order only in association).
//
// Roofline: HBM.  Algorithmic bytes = 12*nnz + 4*(nrow+1) + 8*nvar*(ncol_touched + nrow)   (SURVEY 8d).
#include "scan.cuh"
#include "types.cuh"

namespace ib {

constexpr int APPLY_CHUNK = 32;

// ---- plan: (row, begin, end, partial slot | -1) items ---------------------------------------------------
struct PlanLoad {
    const int32_t *rowptr;
    __device__ uint4 operator()(int64_t r) const
    {
        int len = rowptr[r + 1] - rowptr[r];
        uint32_t nch = len <= APPLY_CHUNK ? 1u : (uint32_t)((len + APPLY_CHUNK - 1) / APPLY_CHUNK);
        return make_uint4(nch, nch > 1 ? nch : 0u, nch > 1 ? 1u : 0u, 0u);
    }
};
struct PlanApply {
    const int32_t *rowptr;
    int4 *items;
    int4 *split_rows;
    __device__ void operator()(int64_t r, uint4 prefix, uint4 e) const
    {
        int b = rowptr[r], end = rowptr[r + 1];
        if (e.x == 1) {
            items[prefix.x] = make_int4((int)r, b, end, -1);
        } else {
            for (uint32_t q = 0; q < e.x; ++q) {
                int lo = b + (int)q * APPLY_CHUNK;
                items[prefix.x + q] = make_int4((int)r, lo, min(end, lo + APPLY_CHUNK), (int)(prefix.y + q));
            }
            split_rows[prefix.z] = make_int4((int)r, (int)prefix.y, (int)e.x, 0);
        }
    }
};

static void make_plan(Weighted &W, cudaStream_t s)
{
    ApplyPlan &P = W.plan;
    if (P.ready) return;
    if (W.nrow > 0) {
        ScanState<uint4> st;
        const uint4 z = make_uint4(0, 0, 0, 0);
        transform_scan<uint4>(st, W.nrow, PlanLoad{W.rowptr.p}, AddU32x4{}, z, s);
        uint4 tot;
        IB_CUDA(cudaMemcpyAsync(&tot, st.total.p, sizeof tot, cudaMemcpyDeviceToHost, s));
        IB_CUDA(cudaStreamSynchronize(s));
        P.nitems = tot.x; P.npartials = tot.y; P.nsplit_rows = tot.z;
        P.items.alloc((size_t)P.nitems, s);
        P.split_rows.alloc((size_t)P.nsplit_rows, s);
        scan_apply<uint4>(st, PlanLoad{W.rowptr.p}, AddU32x4{}, z, PlanApply{W.rowptr.p, P.items.p, P.split_rows.p}, s);
    }
    P.ready = true;
}

// ---- main kernel ----------------------------------------------------------------------------------------
// A group of L lanes (L = 1..32, power of two) owns one (item, field-tile) unit; lane fl of the group owns VEC
// consecutive fields.  No cross-lane traffic: every lane walks the item's nonzeros and keeps its own sums.
template <int VEC>
__global__ void __launch_bounds__(256) spmm_kernel(const int4 *__restrict__ items, int64_t nunits, int ntile, int L,
                                                   const int32_t *__restrict__ col, const double *__restrict__ val,
                                                   const double *__restrict__ X, int64_t ldx, int nvar,
                                                   double *__restrict__ Y, int64_t ldy, double *__restrict__ partials,
                                                   int64_t ldp, const double *__restrict__ factor)
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
        double acc[VEC];
#pragma unroll
        for (int k = 0; k < VEC; ++k) acc[k] = 0.0;
        int j = it.y;
        for (; j + 4 <= it.z; j += 4) {
            int c[4];
            double v[4];
            double x[4][VEC];
#pragma unroll
            for (int q = 0; q < 4; ++q) { c[q] = __ldg(col + j + q); v[q] = __ldg(val + j + q); }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const double *xp = X + (int64_t)c[q] * ldx + f0;
                if (VEC == 2) { double2 t = __ldg(reinterpret_cast<const double2 *>(xp)); x[q][0] = t.x; x[q][VEC - 1] = t.y; }
                else x[q][0] = __ldg(xp);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int k = 0; k < VEC; ++k) acc[k] = acc[k] + v[q] * x[q][k];
        }
        for (; j < it.z; ++j) {
            const int c = __ldg(col + j);
            const double v = __ldg(val + j);
            const double *xp = X + (int64_t)c * ldx + f0;
            if (VEC == 2) { double2 t = __ldg(reinterpret_cast<const double2 *>(xp)); acc[0] = acc[0] + v * t.x; acc[VEC - 1] = acc[VEC - 1] + v * t.y; }
            else acc[0] = acc[0] + v * __ldg(xp);
        }
        double *out;
        if (it.w < 0) {
            out = Y + (int64_t)it.x * ldy + f0;
            if (factor) {
#pragma unroll
                for (int k = 0; k < VEC; ++k) acc[k] = acc[k] * factor[f0 + k];
            }
        } else {
            out = partials + (int64_t)it.w * ldp + f0;
        }
        if (VEC == 2) *reinterpret_cast<double2 *>(out) = make_double2(acc[0], acc[VEC - 1]);
        else out[0] = acc[0];
    }
}

__global__ void spmm_fixup_kernel(const int4 *__restrict__ split_rows, int64_t nsplit, const double *__restrict__ partials,
                                  int64_t ldp, int nvar, double *__restrict__ Y, int64_t ldy, const double *__restrict__ factor)
{
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t r = t / nvar;
    int f = (int)(t - r * nvar);
    if (r >= nsplit) return;
    int4 sr = split_rows[r];
    double s = partials[(int64_t)sr.y * ldp + f];
    for (int q = 1; q < sr.z; ++q) s = s + partials[(int64_t)(sr.y + q) * ldp + f];
    if (factor) s = s * factor[f];
    Y[(int64_t)sr.x * ldy + f] = s;
}

// ---- conservation correction (ibmisc Weighted_Eigen::apply_e; matrix_formats.rst:156-165) ---------------
// out[b][f] = sum over rows r of block b's slice of w[r] * A[r][f]
__global__ void __launch_bounds__(256) wdot_partial_kernel(const double *__restrict__ w, const double *__restrict__ A, int64_t lda,
                                                           int64_t nrow, int nvar, int64_t rows_per_block, double *__restrict__ out)
{
    int64_t r0 = (int64_t)blockIdx.x * rows_per_block, r1 = min(nrow, r0 + rows_per_block);
    for (int f = threadIdx.x; f < nvar; f += blockDim.x) {
        double s = 0;
        for (int64_t r = r0; r < r1; ++r) s = s + w[r] * A[r * lda + f];
        out[(int64_t)blockIdx.x * nvar + f] = s;
    }
}
__global__ void wdot_final_kernel(const double *__restrict__ pa, int nba, const double *__restrict__ pb, int nbb, int nvar,
                                  double *__restrict__ factor)
{
    int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nvar) return;
    double ta = 0, tb = 0;
    for (int b = 0; b < nba; ++b) ta = ta + pa[(int64_t)b * nvar + f];
    for (int b = 0; b < nbb; ++b) tb = tb + pb[(int64_t)b * nvar + f];
    factor[f] = ta * (1. / tb);
}
__global__ void scale_fields_kernel(double *__restrict__ Y, int64_t ldy, int64_t nrow, int nvar, const double *__restrict__ factor)
{
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t r = t / nvar;
    int f = (int)(t - r * nvar);
    if (r < nrow) Y[r * ldy + f] = Y[r * ldy + f] * factor[f];
}

static void wdot(const double *w, const double *A, int64_t lda, int64_t nrow, int nvar, DBuf<double> &part, int &nb, cudaStream_t s)
{
    nb = (int)std::min<int64_t>(std::max<int64_t>(1, cdiv(nrow, 64)), 4 * sm_count());
    int64_t rpb = cdiv(std::max<int64_t>(nrow, 1), nb);
    nb = (int)cdiv(std::max<int64_t>(nrow, 1), rpb);
    part.alloc((size_t)nb * nvar, s);
    wdot_partial_kernel<<<nb, 256, 0, s>>>(w, A, lda, nrow, nvar, rpb, part.p);
    IB_LAUNCHED();
}

void apply_dev(Weighted &W, const double *dX, int64_t ldx, int nvar, double *dY, int64_t ldy, bool force_conservation,
               cudaStream_t s)
{
    ensure_device();
    if (nvar <= 0 || W.nrow == 0) return;
    if (ldx < nvar || ldy < nvar) fail(IB_ERR_ARG, "apply: leading dimension smaller than nvar");
    make_plan(W, s);
    ApplyPlan &P = W.plan;
    const bool vec2 = (nvar % 2 == 0) && (ldx % 2 == 0) && (ldy % 2 == 0) && ((uintptr_t)dX % 16 == 0) && ((uintptr_t)dY % 16 == 0);
    const int VEC = vec2 ? 2 : 1;
    const int lanes_needed = (nvar + VEC - 1) / VEC;
    int L = 1;
    while (L < 32 && L < lanes_needed) L <<= 1;
    const int ntile = (lanes_needed + L - 1) / L;
    const int64_t nunits = P.nitems * ntile;
    const int64_t ldp = (nvar + 1) & ~1;
    DBuf<double> partials((size_t)P.npartials * ldp, s);

    const int G = 32 / L;
    int64_t warps_needed = cdiv(nunits, G);
    int64_t blocks = std::min<int64_t>(cdiv(warps_needed, 8), (int64_t)sm_count() * 8);
    if (blocks < 1) blocks = 1;
    if (VEC == 2)
        spmm_kernel<2><<<(unsigned)blocks, 256, 0, s>>>(P.items.p, nunits, ntile, L, W.col.p, W.val.p, dX, ldx, nvar, dY, ldy, partials.p, ldp, nullptr);
    else
        spmm_kernel<1><<<(unsigned)blocks, 256, 0, s>>>(P.items.p, nunits, ntile, L, W.col.p, W.val.p, dX, ldx, nvar, dY, ldy, partials.p, ldp, nullptr);
    IB_LAUNCHED();
    if (P.nsplit_rows > 0) {
        int64_t tot = P.nsplit_rows * nvar;
        spmm_fixup_kernel<<<(unsigned)cdiv(tot, 256), 256, 0, s>>>(P.split_rows.p, P.nsplit_rows, partials.p, ldp, nvar, dY, ldy, nullptr);
        IB_LAUNCHED();
    }
    if (!W.conservative && force_conservation) {
        DBuf<double> pa, pb, factor((size_t)nvar, s);
        int nba, nbb;
        wdot(W.Mw.p, dX, ldx, W.ncol, nvar, pa, nba, s);
        wdot(W.wM.p, dY, ldy, W.nrow, nvar, pb, nbb, s);
        wdot_final_kernel<<<(unsigned)cdiv(nvar, 128), 128, 0, s>>>(pa.p, nba, pb.p, nbb, nvar, factor.p);
        IB_LAUNCHED();
        int64_t tot = W.nrow * (int64_t)nvar;
        scale_fields_kernel<<<(unsigned)cdiv(tot, 256), 256, 0, s>>>(dY, ldy, W.nrow, nvar, factor.p);
        IB_LAUNCHED();
    }
}

}   // namespace ib
