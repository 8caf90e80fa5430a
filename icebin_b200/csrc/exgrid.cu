// exgrid.cu -- exchange-grid generation on the GPU for a rectilinear ice grid (the step in front of
// the path).
//
// Replaces gridgen/GridGen_Exchange.cpp:175-284 (make_exchange_grid + overlap_callback: RTree search of the ice cells
// under each GCM cell's bounding box, CGAL polygon intersection, area of the overlap polygon) for the case the reference
// runs in production: GCM cells are polygons already projected into the ice grid's plane (lon-lat cells whose sides were
// subdivided before projection, so they may be slightly non-convex), ice cells are the rectangles of a rectilinear grid.
// The overlap of a polygon with an axis-aligned rectangle is Sutherland-Hodgman clipping against four half planes; its
// area is the shoelace sum (Cell::proj_area, Grid.cpp).  One thread per (GCM polygon, ice cell of its bounding box):
//   pass 1  clip -> area per candidate            pass 2  scan of (area > 0) -> compacted (iA, iI, area) triples
// Output order: GCM polygons in input order (the reference walks gridA the same way), ice cells ascending inside one
// (the reference's RTree order is unspecified).  Compiled with -fmad=false so that the CPU oracle's restatement of the
// same arithmetic is bit-identical.
#include <algorithm>

#include "scan.cuh"
#include "types.cuh"

namespace ib {

constexpr int EX_MAXV = 96;       // vertices of a projected GCM cell (4 sides x subdivisions) + 4 for the clip

struct ExArgs {
    const int32_t *pptr;          // [npoly+1] vertex ranges
    const double *pxy;            // [nv][2]
    const int64_t *cand_ptr;      // [npoly+1] candidate ranges
    const int32_t *bs0, *bnf, *bf0;       // per polygon's bounding box: first index on the slow axis, count and first index on the fast axis
    int major_x;                  // 1: iI = ix*ny + iy (x slow; PISM order, searise_grid.cpp:104-108)   0: iI = iy*nx + ix
    const double *xe, *ye;        // ice cell edges [nx+1], [ny+1]
    int nx, ny;
    int npoly;
};

// Clip polygon (n vertices in ax/ay) against one half plane; returns the vertex count written to (bx, by).
// side 0: x >= c   1: x <= c   2: y >= c   3: y <= c
__device__ __host__ inline int ex_clip_edge(const double *ax, const double *ay, int n, int side, double c, double *bx, double *by)
{
    int m = 0;
    for (int i = 0; i < n; ++i) {
        const int j = (i + 1 == n) ? 0 : i + 1;
        const double px = ax[i], py = ay[i], qx = ax[j], qy = ay[j];
        const double pv = (side < 2) ? px : py, qv = (side < 2) ? qx : qy;
        const bool pin = (side & 1) ? (pv <= c) : (pv >= c);
        const bool qin = (side & 1) ? (qv <= c) : (qv >= c);
        if (m + 2 > EX_MAXV) break;        // pathological (strongly non-convex) input: never write past the buffers
        if (pin) { bx[m] = px; by[m] = py; ++m; }
        if (pin != qin) {
            const double t = (c - pv) / (qv - pv);
            if (side < 2) { bx[m] = c; by[m] = py + t * (qy - py); }
            else { bx[m] = px + t * (qx - px); by[m] = c; }
            ++m;
        }
    }
    return m;
}
// area of polygon (ax, ay)[0..n) clipped to [xl, xh] x [yl, yh]
__device__ __host__ inline double ex_clip_area(const double *px, const double *py, int n, double xl, double xh, double yl, double yh)
{
    double ax[EX_MAXV], ay[EX_MAXV], bx[EX_MAXV], by[EX_MAXV];
    for (int i = 0; i < n; ++i) { ax[i] = px[i]; ay[i] = py[i]; }
    int m = ex_clip_edge(ax, ay, n, 0, xl, bx, by);
    if (m < 3) return 0.0;
    m = ex_clip_edge(bx, by, m, 1, xh, ax, ay);
    if (m < 3) return 0.0;
    m = ex_clip_edge(ax, ay, m, 2, yl, bx, by);
    if (m < 3) return 0.0;
    m = ex_clip_edge(bx, by, m, 3, yh, ax, ay);
    if (m < 3) return 0.0;
    double s = 0.0;
    for (int i = 0; i < m; ++i) {
        const int j = (i + 1 == m) ? 0 : i + 1;
        s = s + (ax[i] * ay[j] - ax[j] * ay[i]);
    }
    s = 0.5 * s;
    return s < 0 ? -s : s;
}

__global__ void __launch_bounds__(128) ex_area_kernel(ExArgs a, int64_t ncand, double *area)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ncand) return;
    int lo = 0, hi = a.npoly;              // polygon owning candidate t: last p with cand_ptr[p] <= t
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (a.cand_ptr[mid] <= t) lo = mid; else hi = mid;
    }
    const int p = lo;
    const int64_t k = t - a.cand_ptr[p];
    const int nf = a.bnf[p];
    const int is = a.bs0[p] + (int)(k / nf), ifa = a.bf0[p] + (int)(k % nf);
    const int ix = a.major_x ? is : ifa, iy = a.major_x ? ifa : is;
    const int vb = a.pptr[p], n = a.pptr[p + 1] - vb;
    double px[EX_MAXV - 4], py[EX_MAXV - 4];
    for (int i = 0; i < n; ++i) { px[i] = a.pxy[2 * (vb + i)]; py[i] = a.pxy[2 * (vb + i) + 1]; }
    area[t] = ex_clip_area(px, py, n, a.xe[ix], a.xe[ix + 1], a.ye[iy], a.ye[iy + 1]);
}

struct ExFlagLoad {
    const double *area;
    __device__ uint32_t operator()(int64_t t) const { return area[t] > 0.0 ? 1u : 0u; }
};
struct ExEmit {
    ExArgs a;
    const double *area;
    const int64_t *iA;            // sparse GCM index of each polygon
    int32_t *ex_ij;
    double *ex_area;
    __device__ void operator()(int64_t t, uint32_t prefix, uint32_t flag) const
    {
        if (!flag) return;
        int lo = 0, hi = a.npoly;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (a.cand_ptr[mid] <= t) lo = mid; else hi = mid;
        }
        const int p = lo;
        const int64_t k = t - a.cand_ptr[p];
        const int nf = a.bnf[p];
        const int is = a.bs0[p] + (int)(k / nf), ifa = a.bf0[p] + (int)(k % nf);
        ex_ij[2 * (int64_t)prefix] = (int32_t)iA[p];
        ex_ij[2 * (int64_t)prefix + 1] = a.major_x ? is * a.ny + ifa : is * a.nx + ifa;
        ex_area[prefix] = area[t];
    }
};

// Host driver.  Returns the number of exchange cells; with ex_ij == nullptr only counts.  The result of the last
// count call is kept (per thread) so that the second call only copies.
int64_t exgrid_rectilinear(int npoly, const int32_t *pptr, const double *pxy, const int64_t *iA, int nx, const double *xe, int ny,
                           const double *ye, int iI_major_x, int32_t *ex_ij, double *ex_area, cudaStream_t s)
{
    static thread_local std::vector<int32_t> r_ij;
    static thread_local std::vector<double> r_area;
    if (ex_ij) {
        if (!ex_area) fail(IB_ERR_ARG, "exgrid: ex_area is NULL");
        std::copy(r_ij.begin(), r_ij.end(), ex_ij);
        std::copy(r_area.begin(), r_area.end(), ex_area);
        return (int64_t)r_area.size();
    }
    ensure_device();
    r_ij.clear(); r_area.clear();
    if (npoly < 0 || nx <= 0 || ny <= 0 || !xe || !ye || (npoly > 0 && (!pptr || !pxy || !iA))) fail(IB_ERR_ARG, "exgrid: bad argument");
    if ((int64_t)nx * ny >= (int64_t(1) << 31)) fail(IB_ERR_ARG, "exgrid: ice grid too large for int32 cell ids");
    for (int i = 0; i < nx; ++i) if (!(xe[i + 1] > xe[i])) fail(IB_ERR_ARG, "exgrid: x edges must be strictly increasing");
    for (int i = 0; i < ny; ++i) if (!(ye[i + 1] > ye[i])) fail(IB_ERR_ARG, "exgrid: y edges must be strictly increasing");
    if (npoly == 0) return 0;
    // bounding boxes -> candidate ice cells (the RTree search of the reference, GridGen_Exchange.cpp:266-272)
    std::vector<int32_t> bs0((size_t)npoly), bnf((size_t)npoly), bf0((size_t)npoly);
    std::vector<int64_t> cand((size_t)npoly + 1, 0);
    for (int p = 0; p < npoly; ++p) {
        const int vb = pptr[p], n = pptr[p + 1] - vb;
        if (n < 3 || n > EX_MAXV - 4) fail(IB_ERR_ARG, "exgrid: polygon %d has %d vertices (3..%d supported)", p, n, EX_MAXV - 4);
        if (iA[p] < 0 || iA[p] >= (int64_t(1) << 31)) fail(IB_ERR_ARG, "exgrid: GCM index out of the int32 range");
        double x0 = pxy[2 * vb], x1 = x0, y0 = pxy[2 * vb + 1], y1 = y0;
        for (int i = 1; i < n; ++i) {
            x0 = std::min(x0, pxy[2 * (vb + i)]); x1 = std::max(x1, pxy[2 * (vb + i)]);
            y0 = std::min(y0, pxy[2 * (vb + i) + 1]); y1 = std::max(y1, pxy[2 * (vb + i) + 1]);
        }
        // cells whose interval meets [lo, hi]: from the cell holding lo to the last cell starting before hi
        auto range = [](const double *e, int ncell, double lo, double hi, int32_t &first, int32_t &cnt) {
            int a = (int)(std::upper_bound(e, e + ncell + 1, lo) - e) - 1;
            int b = (int)(std::lower_bound(e, e + ncell + 1, hi) - e);
            a = std::max(a, 0); b = std::min(b, ncell);
            first = a; cnt = std::max(0, b - a);
        };
        int32_t fx, cx, fy, cy;
        range(xe, nx, x0, x1, fx, cx);
        range(ye, ny, y0, y1, fy, cy);
        bs0[p] = iI_major_x ? fx : fy;
        bf0[p] = iI_major_x ? fy : fx;
        bnf[p] = std::max(1, iI_major_x ? cy : cx);       // no candidates: keep the divisor harmless
        cand[p + 1] = cand[p] + (int64_t)cx * cy;
    }
    const int64_t ncand = cand[npoly];
    if (ncand == 0) return 0;
    if (ncand >= (int64_t(1) << 32)) fail(IB_ERR_ARG, "exgrid: too many candidate cells");
    const int nv = pptr[npoly];
    DBuf<int32_t> d_pptr((size_t)npoly + 1, s), d_bs0((size_t)npoly, s), d_bnf((size_t)npoly, s), d_bf0((size_t)npoly, s);
    DBuf<double> d_pxy((size_t)2 * nv, s), d_xe((size_t)nx + 1, s), d_ye((size_t)ny + 1, s), d_area((size_t)ncand, s);
    DBuf<int64_t> d_cand((size_t)npoly + 1, s), d_iA((size_t)npoly, s);
    d_pptr.upload(pptr, (size_t)npoly + 1); d_bs0.upload(bs0.data(), (size_t)npoly); d_bnf.upload(bnf.data(), (size_t)npoly);
    d_bf0.upload(bf0.data(), (size_t)npoly); d_pxy.upload(pxy, (size_t)2 * nv); d_xe.upload(xe, (size_t)nx + 1);
    d_ye.upload(ye, (size_t)ny + 1); d_cand.upload(cand.data(), (size_t)npoly + 1); d_iA.upload(iA, (size_t)npoly);
    ExArgs a{d_pptr.p, d_pxy.p, d_cand.p, d_bs0.p, d_bnf.p, d_bf0.p, iI_major_x, d_xe.p, d_ye.p, nx, ny, npoly};
    ex_area_kernel<<<(unsigned)cdiv(ncand, 128), 128, 0, s>>>(a, ncand, d_area.p);
    IB_LAUNCHED();
    ScanState<uint32_t> st;
    transform_scan<uint32_t>(st, ncand, ExFlagLoad{d_area.p}, AddU32{}, 0u, s);
    uint32_t nex = 0;
    IB_CUDA(cudaMemcpyAsync(&nex, st.total.p, sizeof nex, cudaMemcpyDeviceToHost, s));
    IB_CUDA(cudaStreamSynchronize(s));
    if (nex == 0) return 0;
    DBuf<int32_t> d_ij((size_t)2 * nex, s);
    DBuf<double> d_ea((size_t)nex, s);
    scan_apply<uint32_t>(st, ExFlagLoad{d_area.p}, AddU32{}, 0u, ExEmit{a, d_area.p, d_iA.p, d_ij.p, d_ea.p}, s);
    r_ij.resize((size_t)2 * nex); r_area.resize((size_t)nex);
    d_ij.download(r_ij.data(), (size_t)2 * nex);
    d_ea.download(r_area.data(), (size_t)nex);
    IB_CUDA(cudaStreamSynchronize(s));
    return (int64_t)nex;
}

}   // namespace ib
