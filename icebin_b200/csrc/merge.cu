// merge.cu -- device-resident multi-sheet merge on the A/E grid: VectorMultivec::to_dense_scale + to_dense
// (slib/icebin/multivec.cpp:35-81) for contributions that stay in HBM.
//
// Every ice sheet (one per GPU) produces UNSCALED sums num[row][field] = (M_unscaled x)[row] and weights wM[row]
// for its own dense A/E rows (IceCoupler.cpp:445-457).  The merged value of a cell is sum_sheets(num)/sum_sheets(w).
// Only cells touched by more than one sheet need the exchange: their rows are packed into a contiguous buffer,
// summed across ranks (NCCL all-reduce over NVLink, issued by the caller), and unpacked with the division fused;
// all other rows are divided by their local weight in the same kernel.
#include "radix_sort.cuh"
#include "types.cuh"
#include "util.cuh"

namespace ib {

// buf[k][f] = loc[k] >= 0 ? Y[loc[k]][f] : 0
__global__ void merge_pack_kernel(const double *__restrict__ Y, int64_t ldy, int nvar, const int32_t *__restrict__ loc, int64_t nshared,
                                  double *__restrict__ buf)
{
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t k = t / nvar;
    int f = (int)(t - k * nvar);
    if (k >= nshared) return;
    int32_t r = loc[k];
    buf[k * nvar + f] = r >= 0 ? Y[(int64_t)r * ldy + f] : 0.0;
}

// Y[r][f] = shared_slot[r] >= 0 ? buf[slot][f] * (1/wtot[slot]) : Y[r][f] * (1/wlocal[r])
__global__ void merge_unpack_kernel(double *__restrict__ Y, int64_t ldy, int nvar, int64_t nrow, const int32_t *__restrict__ slot_of_row,
                                    const double *__restrict__ buf, const double *__restrict__ wtot, const double *__restrict__ wlocal)
{
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t r = t / nvar;
    int f = (int)(t - r * nvar);
    if (r >= nrow) return;
    int32_t k = slot_of_row[r];
    double *y = Y + r * ldy + f;
    if (k >= 0) *y = buf[(int64_t)k * nvar + f] * (1. / wtot[k]);     // to_dense: val * scaleE, scaleE = 1/sum(weights)
    else *y = *y * (1. / wlocal[r]);
}

void merge_pack(const double *Y, int64_t ldy, int nvar, const int32_t *loc, int64_t nshared, double *buf, cudaStream_t s)
{
    if (nshared <= 0 || nvar <= 0) return;
    merge_pack_kernel<<<(unsigned)cdiv(nshared * nvar, 256), 256, 0, s>>>(Y, ldy, nvar, loc, nshared, buf);
    IB_LAUNCHED();
}
void merge_unpack(double *Y, int64_t ldy, int nvar, int64_t nrow, const int32_t *slot_of_row, const double *buf, const double *wtot,
                  const double *wlocal, cudaStream_t s)
{
    if (nrow <= 0 || nvar <= 0) return;
    merge_unpack_kernel<<<(unsigned)cdiv(nrow * nvar, 256), 256, 0, s>>>(Y, ldy, nvar, nrow, slot_of_row, buf, wtot, wlocal);
    IB_LAUNCHED();
}

// ---- VectorMultivec::to_dense_scale + to_dense for host callers (multivec.cpp:35-81), computed on the device --------
// The reference walks the entries in input order: scaleE(iE) += w_i, then denseE(iE) (+)= val_i * scaleE(iE).  Here the
// entries are sorted by cell (stable: input order inside a cell) and one thread per (cell, variable) repeats exactly
// those sums in that order (mul then add, compiled with -fmad=false): bit-identical to the sequential loop.
__global__ void mv_key_kernel(const int64_t *index, int64_t n, int64_t nE, uint32_t *key, uint32_t *perm, int *bad)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t e = index[i];
    if (e < 0 || e >= nE) { atomicExch(bad, 1); key[i] = 0; }
    else key[i] = (uint32_t)e;
    perm[i] = (uint32_t)i;
}
__global__ void mv_scale_kernel(const int32_t *ptr, const uint32_t *perm, const double *w, int64_t nE, double *scaleE)
{
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nE) return;
    double s = 0;
    for (int32_t j = ptr[e]; j < ptr[e + 1]; ++j) s += w[perm[j]];
    scaleE[e] = 1. / s;
}
__global__ void mv_dense_kernel(const int32_t *ptr, const uint32_t *perm, const double *vals, int nvar, int64_t nE, const double *scaleE,
                                double fill, double *dense)
{
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t e = t % nE;
    int v = (int)(t / nE);
    if (v >= nvar) return;
    const int32_t b = ptr[e], en = ptr[e + 1];
    double d = fill;
    if (en > b) {
        const double sc = scaleE[e];
        d = vals[(int64_t)perm[b] * nvar + v] * sc;
        for (int32_t j = b + 1; j < en; ++j) d = d + vals[(int64_t)perm[j] * nvar + v] * sc;
    }
    dense[(int64_t)v * nE + e] = d;
}
__global__ void mv_u32_to_i32_kernel(const uint32_t *a, int64_t n, int32_t *b)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) b[i] = (int32_t)a[i];
}

void multivec_to_dense(int64_t n, const int64_t *index, const double *weights, const double *vals, int nvar, int64_t nE, double fill,
                       double *scaleE, double *dense, cudaStream_t s)
{
    ensure_device();
    if (nE <= 0) return;
    if (nE >= (int64_t(1) << 31) || n >= (int64_t(1) << 31)) fail(IB_ERR_ARG, "multivec_to_dense: more than 2^31 cells or entries");
    DBuf<int64_t> dindex((size_t)n, s);
    DBuf<double> dw((size_t)n, s), dvals((size_t)n * nvar, s), dscale((size_t)nE, s), ddense((size_t)nE * nvar, s);
    DBuf<uint32_t> k0((size_t)n, s), k1((size_t)n, s), p0((size_t)n, s), p1((size_t)n, s);
    DBuf<int32_t> sorted((size_t)n, s), ptr((size_t)nE + 1, s);
    DBuf<int> bad(1, s);
    bad.zero();
    ptr.zero();
    dindex.upload(index, (size_t)n); dw.upload(weights, (size_t)n); dvals.upload(vals, (size_t)n * nvar);
    if (n > 0) {
        mv_key_kernel<<<(unsigned)cdiv(n, 256), 256, 0, s>>>(dindex.p, n, nE, k0.p, p0.p, bad.p);
        IB_LAUNCHED();
        const int w = radix_sort_pairs<uint32_t, uint32_t>(k0.p, p0.p, k1.p, p1.p, n, bits_for(nE), s);
        mv_u32_to_i32_kernel<<<(unsigned)cdiv(n, 256), 256, 0, s>>>(w ? k1.p : k0.p, n, sorted.p);
        IB_LAUNCHED();
        seg_ptr_kernel<<<(unsigned)cdiv(n, 256), 256, 0, s>>>(sorted.p, n, nE, ptr.p);
        IB_LAUNCHED();
        const uint32_t *perm = w ? p1.p : p0.p;
        mv_scale_kernel<<<(unsigned)cdiv(nE, 256), 256, 0, s>>>(ptr.p, perm, dw.p, nE, dscale.p);
        IB_LAUNCHED();
        mv_dense_kernel<<<(unsigned)cdiv(nE * nvar, 256), 256, 0, s>>>(ptr.p, perm, dvals.p, nvar, nE, dscale.p, fill, ddense.p);
        IB_LAUNCHED();
    } else {
        mv_scale_kernel<<<(unsigned)cdiv(nE, 256), 256, 0, s>>>(ptr.p, nullptr, dw.p, nE, dscale.p);
        IB_LAUNCHED();
        mv_dense_kernel<<<(unsigned)cdiv(nE * nvar, 256), 256, 0, s>>>(ptr.p, nullptr, dvals.p, nvar, nE, dscale.p, fill, ddense.p);
        IB_LAUNCHED();
    }
    int hbad = 0;
    bad.download(&hbad, 1);
    dscale.download(scaleE, (size_t)nE);
    ddense.download(dense, (size_t)nE * nvar);
    IB_CUDA(cudaStreamSynchronize(s));
    if (hbad) fail(IB_ERR_ARG, "Index out of range (nE = %ld)", (long)nE);
}

}   // namespace ib
