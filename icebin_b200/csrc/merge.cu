// merge.cu -- device-resident multi-sheet merge on the A/E grid: VectorMultivec::to_dense_scale + to_dense
// (slib/icebin/multivec.cpp:35-81) for contributions that stay in HBM.
//
// Every ice sheet (one per GPU) produces UNSCALED sums num[row][field] = (M_unscaled x)[row] and weights wM[row]
// for its own dense A/E rows (IceCoupler.cpp:445-457).  The merged value of a cell is sum_sheets(num)/sum_sheets(w).
// Only cells touched by more than one sheet need the exchange: their rows are packed into a contiguous buffer,
// summed across ranks (NCCL all-reduce over NVLink, issued by the caller), and unpacked with the division fused;
// all other rows are divided by their local weight in the same kernel.
#include "types.cuh"

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

}   // namespace ib
