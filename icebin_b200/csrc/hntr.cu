// hntr.cu -- ModelE "Hntr" conservative lat-lon -> lat-lon regrid as a separable pair of kernels.
//
// Replaces icebin::modele::Hntr (slib/icebin/modele/hntr.hpp:63-435, hntr.cpp:63-168; Fortran HNTR4.F90):
//   B(IB,JB) = sum_{JA,IA} F*G*(wtm*WTA+wtb)*A / sum_{JA,IA} F*G*(wtm*WTA+wtb)
// with F the east-west and G the north-south overlap fractions.  Because the weight F*G is a product, the sum is
// evaluated as   sum_IA F(IA,IB) * [ sum_JA G(JA,JB) * w*A ]:
//   stage 1 streams the source once, every thread owning a few CONSECUTIVE source columns (16-byte loads) and
//           accumulating them down the source rows of one destination row -> S[JB][IA] (value and weight);
//   stage 2 contracts S along IA with the F fractions, applies DATMIS, and
//   stage 3 optionally replaces the polar rows by their mean (hntr.hpp:413-434, including the quirk that one
//           missing cell makes the whole polar row missing).
// The partition tables (SINA/SINB/FMIN/.../JMAX) are computed on the HOST with libm exactly as the reference does
// (the equality tests RIA==RIB and SINA==SINB decide the partition; never recompute them with device sin()).
//
// Roofline: HBM.  Algorithmic bytes = sizeof(Src)*nA [+ 8*nA with a weight array] + 8*nB   (SURVEY 8d).
#include <cmath>
#include <cstdlib>

#include "scan.cuh"
#include "types.cuh"

namespace ib {

struct Hntr {
    int imA, jmA, imB, jmB;
    double offiA, dlatA, offiB, dlatB, datmis;
    std::vector<double> sina, sinb, fmin, fmax, gmin, gmax, dxypA, dxypB;
    std::vector<int> imin, imax, jmin, jmax;
    DBuf<double> d_sina, d_fmin, d_fmax, d_gmin, d_gmax, d_dxypB;
    DBuf<int> d_imin, d_imax, d_jmin, d_jmax;
    bool nested_ew = false;     // every B cell is exactly imA/imB >= 8 whole A columns, starting at ib * imA/imB
    // north-south fractions flattened per destination row: G of source row ja for destination row jb is
    // gflat[goff[jb] + ja - jmin[jb]]; gsum[jb] = sum of them (hntr.hpp:215-219)
    std::vector<double> gflat, gsum;
    std::vector<int> goff;
    DBuf<double> d_gflat, d_gsum;
    DBuf<int> d_goff;
};

static std::vector<double> cell_areas(int im, int jm)   // HntrGrid::dxyp, hntr.cpp:34-52 (unit sphere)
{
    std::vector<double> a(jm);
    const double dlon = (2. * M_PI) / im, dlat = M_PI / jm;
    for (int j = 1; j <= jm; ++j) {
        const double south = sin(dlat * (j - jm / 2 - 1));
        const double north = sin(dlat * (j - jm / 2));
        a[j - 1] = dlon * (north - south);
    }
    return a;
}

static void check_spec(int im, int jm, double offi, double dlat)   // HntrSpec ctor heuristic, GridSpec.cpp:125-132
{
    if (im <= 0 || jm <= 0) fail(IB_ERR_ARG, "HntrSpec(%d,%d,...): im and jm must be positive", im, jm);
    const double lim = .1 * (180. * 60. / jm);
    if (dlat < lim)
        fail(IB_ERR_ARG, "dlat in HntrGrid(%d,%d,%f,%f) seems to small; it should be in MINUTES ,not DEGREES: %f", im, jm, offi, dlat, lim);
}

Hntr *hntr_create(int imB, int jmB, double offiB, double dlatB, int imA, int jmA, double offiA, double dlatA, double datmis)
{
    check_spec(imB, jmB, offiB, dlatB);
    check_spec(imA, jmA, offiA, dlatA);
    std::unique_ptr<Hntr> h(new Hntr);
    h->imA = imA; h->jmA = jmA; h->imB = imB; h->jmB = jmB;
    h->offiA = offiA; h->dlatA = dlatA; h->offiB = offiB; h->dlatB = dlatB; h->datmis = datmis;
    h->dxypA = cell_areas(imA, jmA);
    h->dxypB = cell_areas(imB, jmB);
    h->fmin.assign(imB, 0.); h->fmax.assign(imB, 0.); h->imin.assign(imB, 0); h->imax.assign(imB, 0);
    h->gmin.assign(jmB, 0.); h->gmax.assign(jmB, 0.); h->jmin.assign(jmB, 0); h->jmax.assign(jmB, 0);

    // East-west: walk both grids' eastern cell edges in a domain scaled to [0, imA*imB) (hntr.cpp:84-115).
    // The running edge of grid A is ACCUMULATED (+= imB) as upstream does, so the FMAX quotients match bit for bit.
    {
        const double stepA = imB;
        int ia = 1;
        double eastA = (ia + offiA - imA) * imB;
        for (int ib = 1, west_of = imB; ib <= imB; west_of = ib, ++ib) {
            const double eastB = (ib - 1 + offiB) * imA;     // = western edge of B cell ib
            for (; eastA < eastB; eastA += stepA) ++ia;
            h->imax[west_of - 1] = ia;
            if (eastA == eastB) {          // edges coincide: A cell ia belongs wholly to the western B cell
                h->fmax[west_of - 1] = 0;
                ++ia; eastA += stepA;
                h->fmin[ib - 1] = 0;
            } else {                        // A cell ia straddles the edge
                h->fmax[west_of - 1] = (eastA - eastB) / stepA;
                h->fmin[ib - 1] = 1 - h->fmax[west_of - 1];
            }
            h->imin[ib - 1] = ia;
        }
        h->imax[imB - 1] += imA;           // wrap-around of the last B cell
    }
    // North-south in sin(latitude); dlat is in minutes (hntr.cpp:118-168)
    {
        const double rad_per_min = (2. * M_PI) / (360 * 60);
        auto edges = [&](int jm, double dlat) {
            std::vector<double> sn(jm + 1);
            const double eq = .5 * (1 + jm);
            sn[0] = -1;
            for (int j = 1; j < jm; ++j) sn[j] = sin(((j + .5 - eq) * dlat) * rad_per_min);
            sn[jm] = 1;
            return sn;
        };
        h->sina = edges(jmA, dlatA);
        h->sinb = edges(jmB, dlatB);
        h->jmin[0] = 1;
        for (int jb = 1, ja = 1; jb < jmB; ++jb) {
            while (h->sina[ja] < h->sinb[jb]) ++ja;
            h->jmax[jb - 1] = ja;
            if (h->sina[ja] == h->sinb[jb]) {
                ++ja;
            } else {
                h->gmax[jb - 1] = h->sina[ja] - h->sinb[jb];
                h->gmin[jb] = h->sinb[jb] - h->sina[ja - 1];
            }
            h->jmin[jb] = ja;
        }
        h->jmax[jmB - 1] = jmA;
    }
    h->goff.assign(jmB + 1, 0);
    h->gsum.assign(jmB, 0.0);
    for (int jb = 0; jb < jmB; ++jb) {
        h->goff[jb] = (int)h->gflat.size();
        for (int ja = h->jmin[jb]; ja <= h->jmax[jb]; ++ja) {
            double G = h->sina[ja] - h->sina[ja - 1];
            if (ja == h->jmin[jb]) G -= h->gmin[jb];
            if (ja == h->jmax[jb]) G -= h->gmax[jb];
            h->gflat.push_back(G);
            h->gsum[jb] += G;
        }
    }
    h->goff[jmB] = (int)h->gflat.size();
    h->nested_ew = (imA % imB == 0) && (imA % 8 == 0) && (imA / imB >= 8);
    for (int ib = 0; ib < imB && h->nested_ew; ++ib)
        h->nested_ew = h->fmin[ib] == 0 && h->fmax[ib] == 0 && (h->imin[ib] - 1) % imA == ib * (imA / imB) &&
                       h->imax[ib] - h->imin[ib] + 1 == imA / imB;
    return h.release();
}

static void hntr_upload(Hntr &h, cudaStream_t s)
{
    if (h.d_sina.n) return;
    auto up = [&](auto &d, auto const &v) { d.alloc(v.size(), s); d.upload(v.data(), v.size()); };
    up(h.d_sina, h.sina); up(h.d_fmin, h.fmin); up(h.d_fmax, h.fmax); up(h.d_gmin, h.gmin); up(h.d_gmax, h.gmax);
    up(h.d_gflat, h.gflat); up(h.d_gsum, h.gsum); up(h.d_goff, h.goff);
    up(h.d_dxypB, h.dxypB); up(h.d_imin, h.imin); up(h.d_imax, h.imax); up(h.d_jmin, h.jmin); up(h.d_jmax, h.jmax);
    IB_CUDA(cudaStreamSynchronize(s));
}

// ---- stage 1: S[jb][ia] = sum_ja G * w * A[ja][ia] ; T[jb][ia] = sum_ja G * w ------------------------------
template <class SrcT, int CPT>
struct SrcVec;
template <>
struct SrcVec<int16_t, 8> {
    // int16 -> double without the (quarter-rate) I2F.F64 path: 2^52 + (u ^ 0x8000) is exact in the mantissa of a
    // double whose high word is 0x43300000; subtracting 2^52 + 32768 leaves the signed value exactly.
    static __device__ __forceinline__ void load(const int16_t *p, double (&x)[8])
    {
        int4 r = __ldg(reinterpret_cast<const int4 *>(p));
        const unsigned w[4] = {(unsigned)r.x, (unsigned)r.y, (unsigned)r.z, (unsigned)r.w};
        const double magic = 4503599627370496.0 + 32768.0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            x[2 * k] = __hiloint2double(0x43300000, (int)((w[k] & 0xffffu) ^ 0x8000u)) - magic;
            x[2 * k + 1] = __hiloint2double(0x43300000, (int)((w[k] >> 16) ^ 0x8000u)) - magic;
        }
    }
};
template <>
struct SrcVec<double, 2> {
    static __device__ __forceinline__ void load(const double *p, double (&x)[2])
    {
        double2 r = __ldg(reinterpret_cast<const double2 *>(p));
        x[0] = r.x; x[1] = r.y;
    }
};
template <class SrcT>
struct SrcVec<SrcT, 1> {
    static __device__ __forceinline__ void load(const SrcT *p, double (&x)[1]) { x[0] = (double)__ldg(p); }
};

// RS = row split: the RS warps of a block share the same 32*CPT source columns and take every RS-th source row, so
// narrow sources (int16: 8 columns per 16-byte load => few threads per row) still put enough loads in flight; the
// RS partial sums are combined through shared memory in warp order (deterministic).
template <class SrcT, int CPT, bool HASW, int UNR, int RS>
__global__ void __launch_bounds__(128, (sizeof(SrcT) == 2 ? 16 : 8)) hntr_cols_kernel(const SrcT *__restrict__ A, const double *__restrict__ WTA, double wtm, double wtb,
                                                        int imA, const double *__restrict__ sina, const int *__restrict__ jmin,
                                                        const int *__restrict__ jmax, const double *__restrict__ gmin,
                                                        const double *__restrict__ gmax, double *__restrict__ S, double *__restrict__ T)
{
    constexpr int TPR = 128 / RS;                      // threads (column groups) per source row
    __shared__ double red[RS > 1 ? (RS - 1) * TPR * (2 * CPT) : 1];
    const int jb = blockIdx.y;
    const int part = threadIdx.x / TPR, tcol = threadIdx.x % TPR;
    const int ia0 = (blockIdx.x * TPR + tcol) * CPT;
    const bool live = ia0 < imA;
    const int jlo = jmin[jb], jhi = jmax[jb];
    double sv[CPT], sw[CPT];
#pragma unroll
    for (int k = 0; k < CPT; ++k) { sv[k] = 0; sw[k] = 0; }
    const double wconst = wtm * 1.0 + wtb;
    double gsum = 0;
    if (live) {
#pragma unroll UNR
        for (int ja = jlo + part; ja <= jhi; ja += RS) {
            double G = sina[ja] - sina[ja - 1];
            if (ja == jlo) G -= gmin[jb];
            if (ja == jhi) G -= gmax[jb];
            const int64_t off = (int64_t)(ja - 1) * imA + ia0;
            double x[CPT];
            SrcVec<SrcT, CPT>::load(A + off, x);
            if (HASW) {
                double w[CPT];
#pragma unroll
                for (int k = 0; k < CPT; k += (CPT > 1 ? 2 : 1)) {
                    if (CPT > 1) { double2 t = __ldg(reinterpret_cast<const double2 *>(WTA + off + k)); w[k] = t.x; w[k + (CPT > 1)] = t.y; }
                    else w[k] = __ldg(WTA + off + k);
                }
#pragma unroll
                for (int k = 0; k < CPT; ++k) {
                    const double wt = G * (wtm * w[k] + wtb);
                    sw[k] += wt;
                    sv[k] += wt * x[k];
                }
            } else {
                const double wt = G * wconst;
                gsum += wt;
#pragma unroll
                for (int k = 0; k < CPT; ++k) sv[k] += wt * x[k];
            }
        }
    }
    if (RS > 1) {
        if (part > 0) {
            double *r = red + ((part - 1) * TPR + tcol) * (2 * CPT);
#pragma unroll
            for (int k = 0; k < CPT; ++k) { r[k] = sv[k]; r[CPT + k] = HASW ? sw[k] : gsum; }
        }
        __syncthreads();
        if (part > 0) return;
#pragma unroll
        for (int p = 1; p < RS; ++p) {
            const double *r = red + ((p - 1) * TPR + tcol) * (2 * CPT);
#pragma unroll
            for (int k = 0; k < CPT; ++k) { sv[k] += r[k]; if (HASW) sw[k] += r[CPT + k]; }
            if (!HASW) gsum += r[CPT];
        }
    }
    if (!live) return;
    double *so = S + (int64_t)jb * imA + ia0;
    double *to = T + (int64_t)jb * imA + ia0;
#pragma unroll
    for (int k = 0; k < CPT; ++k) {
        so[k] = sv[k];
        to[k] = HASW ? sw[k] : gsum;
    }
}

// ---- stage 2: one warp per destination cell contracts S,T along IA with the F fractions ----------------------
__global__ void __launch_bounds__(256) hntr_rows_kernel(const double *__restrict__ S, const double *__restrict__ T, int imA, int imB, int jmB,
                                                        const int *__restrict__ imin, const int *__restrict__ imax,
                                                        const double *__restrict__ fmin, const double *__restrict__ fmax, double datmis,
                                                        double *__restrict__ B)
{
    const int lane = threadIdx.x & 31;
    const int64_t cell = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (cell >= (int64_t)imB * jmB) return;
    const int jb = (int)(cell / imB), ib = (int)(cell - (int64_t)jb * imB);
    const int lo = imin[ib], hi = imax[ib];
    const double *s = S + (int64_t)jb * imA, *t = T + (int64_t)jb * imA;
    double v = 0, w = 0;
    for (int ir = lo + lane; ir <= hi; ir += 32) {
        const int ia = (ir - 1) % imA;
        double F = 1;
        if (ir == lo) F -= fmin[ib];
        if (ir == hi) F -= fmax[ib];
        v += F * s[ia];
        w += F * t[ia];
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        v += __shfl_xor_sync(0xffffffffu, v, d);
        w += __shfl_xor_sync(0xffffffffu, w, d);
    }
    if (lane == 0) B[cell] = (w == 0) ? datmis : v / w;
}

// ---- stage 3: polar rows -> longitudinal mean (one block per polar row) ---------------------------------------
__global__ void __launch_bounds__(256) hntr_polar_kernel(double *__restrict__ B, int imB, int jmB, double datmis)
{
    __shared__ double sh[256];
    __shared__ int miss;
    const int jb = blockIdx.x == 0 ? 0 : jmB - 1;
    double *row = B + (int64_t)jb * imB;
    if (threadIdx.x == 0) miss = 0;
    __syncthreads();
    double s = 0;
    for (int i = threadIdx.x; i < imB; i += blockDim.x) {
        const double b = row[i];
        if (b == datmis) miss = 1;
        s += b;
    }
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int d = 128; d > 0; d >>= 1) {
        if (threadIdx.x < d) sh[threadIdx.x] += sh[threadIdx.x + d];
        __syncthreads();
    }
    const double mean = miss ? datmis : sh[0] / (double)imB;
    for (int i = threadIdx.x; i < imB; i += blockDim.x) row[i] = mean;
}

// ---- fast path: int16 source, constant weight, east-west exactly nested (ETOPO1 1' -> any ModelE grid) --------
// With F == 1 everywhere the inner sum over a destination cell's source columns is an INTEGER sum, so a thread that
// owns 8 consecutive int16 columns folds each source row with 4 two-way integer dot products (IDP.2A) instead of
// 8 int->fp64 conversions + 8 DFMAs: ~2 instead of ~9.5 instructions per element, which is what lets the kernel run
// at memory speed.  A thread's 8 columns touch at most two destination cells (cell width >= 8): elements k < split
// go to the first (acc0), the rest to the second (acc1).  Integer row sums are exact; only G * sum is rounded.
__device__ __forceinline__ double i32_to_f64(int x)      // exact, without the XU conversion pipe
{
    return __hiloint2double(0x43300000, x ^ 0x80000000) - 4503601774854144.0;      // 2^52 + 2^31
}
template <int UNR>
__global__ void __launch_bounds__(128, 12) hntr_i16_nested_kernel(const int16_t *__restrict__ A, int imA, int cols_per_b,
                                                                  const int *__restrict__ jmin, const int *__restrict__ jmax,
                                                                  const double *__restrict__ gflat, const int *__restrict__ goff,
                                                                  double wconst, double *__restrict__ P)
{
    // blockIdx.z splits the source rows of a destination row into gridDim.z interleaved parts (more CTAs when the
    // destination grid has few rows); part z writes its own slice of P and stage 2 adds the parts in order
    const int jb = blockIdx.y;
    const int nvec = imA >> 3;
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nvec) return;
    const int c0 = v << 3;
    const int split = min(8, (c0 / cols_per_b + 1) * cols_per_b - c0);      // elements of this vector in its first cell
    unsigned m[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) m[q] = (2 * q + 1 < split) ? 0xffffffffu : ((2 * q < split) ? 0x0000ffffu : 0u);
    const int jlo = jmin[jb], jhi = jmax[jb];
    const double *g = gflat + goff[jb] - jlo;
    double acc0 = 0.0, acc1 = 0.0;
    const int4 *src = reinterpret_cast<const int4 *>(A) + v;
#pragma unroll UNR
    for (int ja = jlo + (int)blockIdx.z; ja <= jhi; ja += (int)gridDim.z) {
        const int4 r = __ldg(src + (int64_t)(ja - 1) * nvec);
        const double gw = __ldg(g + ja) * wconst;
        const int w[4] = {r.x, r.y, r.z, r.w};
        int s_all = 0, s_lo = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            s_all = __dp2a_lo(w[q], 0x0101, s_all);
            s_lo = __dp2a_lo((int)((unsigned)w[q] & m[q]), 0x0101, s_lo);
        }
        acc0 += gw * i32_to_f64(s_lo);
        acc1 += gw * i32_to_f64(s_all - s_lo);
    }
    reinterpret_cast<double2 *>(P)[((int64_t)blockIdx.z * gridDim.y + jb) * nvec + v] = make_double2(acc0, acc1);
}

// a group of L lanes (L = 1..32, power of two) per destination cell: add, in a fixed order, the per-vector sums that
// belong to it; weight = cols * w * sum_JA G
__global__ void __launch_bounds__(256) hntr_rows_nested_kernel(const double *__restrict__ P, int nz, int imA, int imB, int jmB, int cols_per_b,
                                                               int L, const double *__restrict__ gsum, double wconst, double datmis,
                                                               double *__restrict__ B)
{
    const int lane = threadIdx.x & 31;
    const int sub = lane % L;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t cell = t / L;
    const int64_t ncell = (int64_t)imB * jmB;
    const bool live = cell < ncell;
    const int jb = live ? (int)(cell / imB) : 0, ib = live ? (int)(cell - (int64_t)jb * imB) : 0;
    const int nvec = imA >> 3;
    const int v_lo = (ib * cols_per_b) >> 3, v_hi = (ib * cols_per_b + cols_per_b - 1) >> 3;
    const int nv = v_hi - v_lo + 1;
    double vsum = 0.0;
    if (live)
        for (int k = sub; k < nv * nz; k += L) {
            const int z = k / nv, v = v_lo + (k - z * nv);
            const double2 tt = reinterpret_cast<const double2 *>(P)[((int64_t)z * jmB + jb) * nvec + v];
            vsum += ((v << 3) / cols_per_b == ib) ? tt.x : tt.y;
        }
    for (int d = L >> 1; d > 0; d >>= 1) vsum += __shfl_xor_sync(0xffffffffu, vsum, d);
    if (!live || sub != 0) return;
    const double wsum = (gsum[jb] * wconst) * (double)cols_per_b;
    B[cell] = (wsum == 0) ? datmis : vsum / wsum;
}

template <class SrcT, int CPT>
static void launch_cols(Hntr &h, const SrcT *A, const double *W, double wtm, double wtb, double *S, double *T, cudaStream_t s)
{
    constexpr int RS = 1;   // row split (measured slower on B200 for the ETOPO1 shapes; kept for narrow grids)
    dim3 grid((unsigned)cdiv(cdiv(h.imA, CPT), 128 / RS), (unsigned)h.jmB);
    if (W)
        hntr_cols_kernel<SrcT, CPT, true, (sizeof(SrcT) == 2 ? 2 : 4), RS><<<grid, 128, 0, s>>>(A, W, wtm, wtb, h.imA, h.d_sina.p, h.d_jmin.p, h.d_jmax.p, h.d_gmin.p, h.d_gmax.p, S, T);
    else
        hntr_cols_kernel<SrcT, CPT, false, (sizeof(SrcT) == 2 ? 2 : 4), RS><<<grid, 128, 0, s>>>(A, W, wtm, wtb, h.imA, h.d_sina.p, h.d_jmin.p, h.d_jmax.p, h.d_gmin.p, h.d_gmax.p, S, T);
    IB_LAUNCHED();
}

void hntr_regrid_dev(Hntr &h, const double *d_wta, double wtm, double wtb, const void *d_src, int src_dtype, double *d_dst,
                     bool mean_polar, cudaStream_t s)
{
    ensure_device();
    hntr_upload(h, s);
    const bool al16 = ((uintptr_t)d_src % 16 == 0) && (d_wta == nullptr || (uintptr_t)d_wta % 16 == 0);
    const int64_t ncell = (int64_t)h.imB * h.jmB;
    if (src_dtype == IB_SRC_I16 && d_wta == nullptr && al16 && h.nested_ew && !g_tune.hntr_generic) {
        const int nvec = h.imA / 8, cpb = h.imA / h.imB;
        const double wconst = wtm * 1.0 + wtb;
        const int64_t xb = cdiv(nvec, 128);
        const int nz = (int)std::min<int64_t>(8, std::max<int64_t>(1, cdiv(3500, xb * h.jmB)));
        DBuf<double> P((size_t)nz * h.jmB * nvec * 2, s);
        dim3 grid((unsigned)xb, (unsigned)h.jmB, (unsigned)nz);
        hntr_i16_nested_kernel<4><<<grid, 128, 0, s>>>((const int16_t *)d_src, h.imA, cpb, h.d_jmin.p, h.d_jmax.p, h.d_gflat.p, h.d_goff.p,
                                                       wconst, P.p);
        IB_LAUNCHED();
        const int work = ((cpb + 7) / 8 + 1) * nz;      // per-vector sums a destination cell adds up
        int L = 1;
        while (L < 32 && 2 * L <= work) L <<= 1;
        hntr_rows_nested_kernel<<<(unsigned)cdiv(ncell * L, 256), 256, 0, s>>>(P.p, nz, h.imA, h.imB, h.jmB, cpb, L, h.d_gsum.p, wconst, h.datmis, d_dst);
        IB_LAUNCHED();
        if (mean_polar) {
            hntr_polar_kernel<<<h.jmB > 1 ? 2 : 1, 256, 0, s>>>(d_dst, h.imB, h.jmB, h.datmis);
            IB_LAUNCHED();
        }
        return;
    }
    DBuf<double> S((size_t)h.jmB * h.imA, s), T((size_t)h.jmB * h.imA, s);
    if (src_dtype == IB_SRC_I16) {
        if (al16 && h.imA % 8 == 0) launch_cols<int16_t, 8>(h, (const int16_t *)d_src, d_wta, wtm, wtb, S.p, T.p, s);
        else launch_cols<int16_t, 1>(h, (const int16_t *)d_src, d_wta, wtm, wtb, S.p, T.p, s);
    } else if (src_dtype == IB_SRC_F64) {
        if (al16 && h.imA % 2 == 0) launch_cols<double, 2>(h, (const double *)d_src, d_wta, wtm, wtb, S.p, T.p, s);
        else launch_cols<double, 1>(h, (const double *)d_src, d_wta, wtm, wtb, S.p, T.p, s);
    } else {
        fail(IB_ERR_ARG, "hntr_regrid: unknown source dtype %d", src_dtype);
    }
    hntr_rows_kernel<<<(unsigned)cdiv(ncell * 32, 256), 256, 0, s>>>(S.p, T.p, h.imA, h.imB, h.jmB, h.d_imin.p, h.d_imax.p, h.d_fmin.p,
                                                                    h.d_fmax.p, h.datmis, d_dst);
    IB_LAUNCHED();
    if (mean_polar) {
        hntr_polar_kernel<<<h.jmB > 1 ? 2 : 1, 256, 0, s>>>(d_dst, h.imB, h.jmB, h.datmis);
        IB_LAUNCHED();
    }
}

// ---- overlap / scaled_regrid_matrix (hntr.hpp:249-339): entries in the reference's emission order -----------
struct OvCount {
    const int *imin, *imax, *jmin, *jmax;
    int imB;
    __device__ uint32_t operator()(int64_t cell) const
    {
        int jb = (int)(cell / imB), ib = (int)(cell - (int64_t)jb * imB);
        return (uint32_t)((jmax[jb] - jmin[jb] + 1) * (imax[ib] - imin[ib] + 1));
    }
};
struct OvEmit {
    const int *imin, *imax, *jmin, *jmax;
    const double *fmin, *fmax, *gmin, *gmax, *sina, *dxypB;
    int imA, imB;
    double R2;   // < 0: scaled_regrid_matrix
    int64_t *rows, *cols;
    double *vals;
    __device__ void operator()(int64_t cell, uint32_t prefix, uint32_t) const
    {
        const int jb = (int)(cell / imB), ib = (int)(cell - (int64_t)jb * imB);
        const int jlo = jmin[jb], jhi = jmax[jb], ilo = imin[ib], ihi = imax[ib];
        double weight = 0;   // sequential, reference order (JA outer, IA inner)
        for (int ja = jlo; ja <= jhi; ++ja) {
            double G = sina[ja] - sina[ja - 1];
            if (ja == jlo) G -= gmin[jb];
            if (ja == jhi) G -= gmax[jb];
            for (int ir = ilo; ir <= ihi; ++ir) {
                double F = 1;
                if (ir == ilo) F -= fmin[ib];
                if (ir == ihi) F -= fmax[ib];
                weight += F * G;
            }
        }
        const double by = 1. / weight;
        int64_t o = prefix;
        for (int ja = jlo; ja <= jhi; ++ja) {
            double G = sina[ja] - sina[ja - 1];
            if (ja == jlo) G -= gmin[jb];
            if (ja == jhi) G -= gmax[jb];
            for (int ir = ilo; ir <= ihi; ++ir, ++o) {
                double F = 1;
                if (ir == ilo) F -= fmin[ib];
                if (ir == ihi) F -= fmax[ib];
                const double fg = F * G;
                rows[o] = cell;
                cols[o] = (int64_t)(ja - 1) * imA + ((ir - 1) % imA);
                vals[o] = (R2 < 0) ? fg * by : fg * by * (R2 * dxypB[jb]);
            }
        }
    }
};

int64_t hntr_matrix(Hntr &h, double R, int64_t *rows, int64_t *cols, double *vals, cudaStream_t s)
{
    ensure_device();
    hntr_upload(h, s);
    const int64_t ncell = (int64_t)h.imB * h.jmB;
    // host-side count first (cheap, exact)
    int64_t total = 0;
    for (int jb = 0; jb < h.jmB; ++jb)
        for (int ib = 0; ib < h.imB; ++ib) total += (int64_t)(h.jmax[jb] - h.jmin[jb] + 1) * (h.imax[ib] - h.imin[ib] + 1);
    if (!rows) return total;
    if (total >= (int64_t(1) << 32)) fail(IB_ERR_ARG, "hntr_matrix: more than 2^32 entries");
    ScanState<uint32_t> st;
    OvCount cnt{h.d_imin.p, h.d_imax.p, h.d_jmin.p, h.d_jmax.p, h.imB};
    transform_scan<uint32_t>(st, ncell, cnt, AddU32{}, 0u, s);
    DBuf<int64_t> dr((size_t)total, s), dc((size_t)total, s);
    DBuf<double> dv((size_t)total, s);
    OvEmit em{h.d_imin.p, h.d_imax.p, h.d_jmin.p, h.d_jmax.p, h.d_fmin.p, h.d_fmax.p, h.d_gmin.p, h.d_gmax.p, h.d_sina.p, h.d_dxypB.p,
              h.imA, h.imB, R < 0 ? -1.0 : R * R, dr.p, dc.p, dv.p};
    scan_apply<uint32_t>(st, cnt, AddU32{}, 0u, em, s);
    dr.download(rows, (size_t)total);
    dc.download(cols, (size_t)total);
    dv.download(vals, (size_t)total);
    IB_CUDA(cudaStreamSynchronize(s));
    return total;
}

void hntr_free(Hntr *h) { delete h; }

void hntr_get(Hntr const &h, int which, double *dout, int32_t *iout)
{
    auto cpd = [&](std::vector<double> const &v) { if (!dout) fail(IB_ERR_ARG, "hntr_get: double buffer required"); std::copy(v.begin(), v.end(), dout); };
    auto cpi = [&](std::vector<int> const &v) { if (!iout) fail(IB_ERR_ARG, "hntr_get: int buffer required"); std::copy(v.begin(), v.end(), iout); };
    switch (which) {
        case 0: cpd(h.sina); break;
        case 1: cpd(h.sinb); break;
        case 2: cpd(h.fmin); break;
        case 3: cpd(h.fmax); break;
        case 4: cpd(h.gmin); break;
        case 5: cpd(h.gmax); break;
        case 6: cpi(h.imin); break;
        case 7: cpi(h.imax); break;
        case 8: cpi(h.jmin); break;
        case 9: cpi(h.jmax); break;
        case 10: cpd(h.dxypA); break;
        case 11: cpd(h.dxypB); break;
        default: fail(IB_ERR_ARG, "hntr_get: unknown selector %d", which);
    }
}
void hntr_dims(Hntr const &h, int *d) { d[0] = h.imA; d[1] = h.jmA; d[2] = h.imB; d[3] = h.jmB; }

}   // namespace ib
