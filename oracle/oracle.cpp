// oracle/oracle.cpp -- TEST INFRASTRUCTURE ONLY (not product code).
//
// CPU restatement of the IceBin regridding hot path, used ONLY by
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg as the
// checker and the reported CPU baseline.  Nothing under icebin_b200/ may call into this file.
//
// The reference (/root/reference, citibeth/icebin) cannot be compiled here: its sparse algebra
// lives in the un-vendored ibmisc/spsparse + Eigen + Blitz++.  This file restates
//   * spsparse::SparseSet / MakeDenseEigen / TupleList  (ibmisc develop, unpinned; semantics inferred
//     from call sites, SURVEY 8c "Inferred ibmisc semantics")
//   * Eigen::SparseMatrix<double,ColMajor,int>: setFromTriplets, sparse*sparse, diag scaling
//   * slib/icebin/IceRegridder_L0.cpp:43-214, IceRegridder.cpp:41-65,
//     RegridMatrices_Dynamic.cpp:19-437, smoother.cpp:8-99, multivec.cpp:35-81,
//     modele/hntr.cpp:21-168, modele/hntr.hpp:204-435
// in the same single-threaded, hash-map + triplet + CSC-product structure as the reference.
//
// PARITY STATUS: the six-matrix path is "parity unpinned" numerically -- the reference holds no
// golden vectors for it, only property tests (tests/test_conserv/test_conserv.py:89-205), which
// tests/test_oracle_properties.py re-runs against this file.  Hntr is pinned by the closed-form
// identities of tests/test_hntr.cpp:332-427,464-494 and the doc example IceRegridder.hpp:150-155.
//
// Build: make -C oracle   (g++ -O2, no dependencies).  C ABI at the bottom (ctypes-friendly).

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <memory>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

namespace orc {

static const double NaN = std::numeric_limits<double>::quiet_NaN();

[[noreturn]] static void fail(const char *fmt, double a = 0, double b = 0, double c = 0)
{
    char buf[512];
    snprintf(buf, sizeof buf, fmt, a, b, c);
    throw std::runtime_error(buf);
}

// ---------------------------------------------------------------------------------------
// spsparse::SparseSet<long,int>: ordered bijection sparse<->dense; add_dense() returns the
// existing id or appends (AbbrGrid.cpp:105-121 relies on ids being handed out sequentially).
struct SparseSet {
    long sparse_extent = -1;
    std::unordered_map<long, int> to_d;
    std::vector<long> to_s;

    int dense_extent() const { return (int)to_s.size(); }
    bool in_sparse(long s) const { return to_d.find(s) != to_d.end(); }
    int add_dense(long s)
    {
        auto it = to_d.find(s);
        if (it != to_d.end()) return it->second;
        int d = (int)to_s.size();
        to_d.emplace(s, d);
        to_s.push_back(s);
        return d;
    }
    int to_dense(long s) const
    {
        auto it = to_d.find(s);
        if (it == to_d.end()) fail("SparseSet::to_dense: sparse index %g missing", (double)s);
        return it->second;
    }
};

// ---------------------------------------------------------------------------------------
// Column-major sparse matrix with sorted inner (row) indices == Eigen::SparseMatrix<double,0,int>
// (eigen_types.hpp:21).
struct Trip { int r, c; double v; };

struct Csc {
    int nrow = 0, ncol = 0;
    std::vector<int> ptr;   // ncol+1
    std::vector<int> row;
    std::vector<double> val;
    int nnz() const { return (int)row.size(); }
};

// Eigen setFromTriplets: duplicates are summed, in input order; inner indices end up sorted.
static Csc assemble(std::vector<Trip> const &t, int nrow, int ncol)
{
    std::vector<int> order(t.size());
    for (size_t i = 0; i < t.size(); ++i) order[i] = (int)i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
        if (t[a].c != t[b].c) return t[a].c < t[b].c;
        return t[a].r < t[b].r;
    });
    Csc M;
    M.nrow = nrow; M.ncol = ncol;
    M.ptr.assign(ncol + 1, 0);
    int last_r = -1, last_c = -1;
    for (int k : order) {
        Trip const &e = t[k];
        if (e.r < 0 || e.r >= nrow || e.c < 0 || e.c >= ncol) fail("assemble: index out of range");
        if (e.r == last_r && e.c == last_c) {
            M.val.back() += e.v;
        } else {
            M.row.push_back(e.r);
            M.val.push_back(e.v);
            M.ptr[e.c + 1]++;
            last_r = e.r; last_c = e.c;
        }
    }
    for (int j = 0; j < ncol; ++j) M.ptr[j + 1] += M.ptr[j];
    return M;
}

// Eigen conservative sparse*sparse (col-major): result(:,j) = sum_k A(:,k)*B(k,j), k ascending.
static Csc mul(Csc const &A, Csc const &B)
{
    if (A.ncol != B.nrow) fail("mul: inner dimensions differ (%g vs %g)", A.ncol, B.nrow);
    Csc C;
    C.nrow = A.nrow; C.ncol = B.ncol;
    C.ptr.assign(B.ncol + 1, 0);
    std::vector<double> acc(A.nrow, 0.0);
    std::vector<char> seen(A.nrow, 0);
    std::vector<int> touched;
    for (int j = 0; j < B.ncol; ++j) {
        touched.clear();
        for (int q = B.ptr[j]; q < B.ptr[j + 1]; ++q) {
            int k = B.row[q];
            double y = B.val[q];
            for (int p = A.ptr[k]; p < A.ptr[k + 1]; ++p) {
                int i = A.row[p];
                double x = A.val[p];
                if (!seen[i]) { seen[i] = 1; acc[i] = x * y; touched.push_back(i); }
                else acc[i] += x * y;
            }
        }
        std::sort(touched.begin(), touched.end());
        for (int i : touched) {
            C.row.push_back(i);
            C.val.push_back(acc[i]);
            seen[i] = 0;
        }
        C.ptr[j + 1] = (int)C.row.size();
    }
    return C;
}

// M * diag(d): scales columns.   diag(d) * M: scales rows.  (map_eigen_diagonal, test_hntr.cpp:271)
static Csc scale_cols(Csc M, std::vector<double> const &d)
{
    if ((int)d.size() != M.ncol) fail("scale_cols: size mismatch");
    for (int j = 0; j < M.ncol; ++j)
        for (int p = M.ptr[j]; p < M.ptr[j + 1]; ++p) M.val[p] = M.val[p] * d[j];
    return M;
}
static Csc scale_rows(std::vector<double> const &d, Csc M)
{
    if ((int)d.size() != M.nrow) fail("scale_rows: size mismatch");
    for (int p = 0; p < M.nnz(); ++p) M.val[p] = d[M.row[p]] * M.val[p];
    return M;
}

// ibmisc sum(M, dim, c): dim 0 -> per-row sums, dim 1 -> per-column sums; '-' -> reciprocal.
// (traversal is storage order.)
static std::vector<double> sum(Csc const &M, int dim, char c)
{
    std::vector<double> s(dim == 0 ? M.nrow : M.ncol, 0.0);
    for (int j = 0; j < M.ncol; ++j)
        for (int p = M.ptr[j]; p < M.ptr[j + 1]; ++p) s[dim == 0 ? M.row[p] : j] += M.val[p];
    if (c == '-') for (double &x : s) x = 1. / x;
    return s;
}
static std::vector<double> invert1(std::vector<double> v)   // RegridMatrices_Dynamic.cpp:19-24
{
    for (double &x : v) x = 1. / x;
    return v;
}

// ---------------------------------------------------------------------------------------
// MakeDenseEigen: run a generator of sparse-indexed tuples through a per-dimension
// sparse->dense transform, collect triplets; assemble when both dims are final.
enum Xf { ADD_DENSE, TO_DENSE_IGNORE_MISSING };
struct STup { long i, j; double v; };   // sparse-indexed tuple as emitted by a generator

static std::vector<Trip> densify(std::vector<STup> const &gen, Xf xf, SparseSet *d0, SparseSet *d1)
{
    std::vector<Trip> out;
    out.reserve(gen.size());
    for (STup const &t : gen) {
        if (xf == ADD_DENSE) {
            int a = d0->add_dense(t.i);
            int b = d1->add_dense(t.j);
            out.push_back({a, b, t.v});
        } else {
            if (!d0->in_sparse(t.i) || !d1->in_sparse(t.j)) continue;
            out.push_back({d0->to_dense(t.i), d1->to_dense(t.j), t.v});
        }
    }
    return out;
}
static Csc to_eigen(std::vector<Trip> t, SparseSet const *d0, SparseSet const *d1, char tr)
{
    if (tr == 'T') {
        for (Trip &e : t) std::swap(e.r, e.c);
        return assemble(t, d1->dense_extent(), d0->dense_extent());
    }
    return assemble(t, d0->dense_extent(), d1->dense_extent());
}

// ---------------------------------------------------------------------------------------
// Grid data (AbbrGrid.hpp:40-147) reduced to what the hot path reads.
struct AGrid {              // AbbrGrid of the GCM 'A' grid
    long nA = 0;            // sparse extent
    std::vector<long> dim;  // realised cells: dense -> sparse
    std::vector<double> native_area;   // dense
};
struct Sheet {              // IceRegridder(_L0) state
    long nI = 0;
    std::vector<long> dimI;             // agridI.dim dense->sparse
    std::vector<double> native_areaI;   // dense
    std::vector<double> centroid_xy;    // dense, [n][2]
    long nX = 0;
    std::vector<int> ex_ij;             // ExchangeGrid::indices, interleaved (iA, iI)
    std::vector<double> ex_area;        // ExchangeGrid::overlaps
    std::vector<double> gridA_proj_area;   // per realised A cell (IceRegridder.cpp:93-119)
    int interp_style = 0;               // 0 = Z_INTERP, 1 = ELEV_CLASS_INTERP
};
struct Gcm {
    AGrid A;
    std::vector<double> hcdefs;
    std::vector<std::unique_ptr<Sheet>> sheets;
    long nhc() const { return (long)hcdefs.size(); }
    long nE() const { return A.nA * nhc(); }
    // indexingHC = Indexing({"A","HC"},{0,0},{nA,nhc},{1,0}) -> HC has the larger stride
    // (pylib/icebin_cython.cpp:56-73, pylib/icebin/ibgrid.py:64-75)
    long iE(long iA, long ihc) const { return ihc * A.nA + iA; }
};

// IceRegridder_L0.cpp:71-93
static void linterp_1d_b(std::vector<double> const &xp, double xx, long idx[2], double w[2])
{
    int n = (int)xp.size();
    int hi = (int)(std::lower_bound(xp.begin(), xp.end(), xx) - xp.begin());
    if (hi <= 0) hi = 1;
    if (hi >= n) fail("Elevation %g out of bounds (%g, %g)", xx, xp.front(), xp.back());
    int lo = hi - 1;
    idx[0] = lo; idx[1] = hi;
    double ratio = (xx - xp[lo]) / (xp[hi] - xp[lo]);
    w[0] = 1.0 - ratio;
    w[1] = ratio;
}
// IceRegridder_L0.cpp:43-67
static int nearest_1d(std::vector<double> const &xp, double xx)
{
    int n = (int)xp.size();
    int hi = (int)(std::lower_bound(xp.begin(), xp.end(), xx) - xp.begin());
    if (hi <= 0) return 0;
    if (hi >= n) return n - 1;
    double dlo = std::abs(xx - xp[hi - 1]);
    double dhi = std::abs(xp[hi] - xx);
    return dlo <= dhi ? hi - 1 : hi;
}

// Generators, gridG == 'X' flavour only (the only one regrid_matrices() binds,
// RegridMatrices_Dynamic.cpp:354-360).
static std::vector<STup> gen_GvEp(Gcm const &g, Sheet const &s, const double *elevmaskI)   // L0.cpp:100-155
{
    if (g.hcdefs.empty()) fail("IceRegridder_L0::GvEp(): hcdefs is zero-length!");
    std::vector<STup> out;
    for (long x = 0; x < s.nX; ++x) {
        long iA = s.ex_ij[2 * x], iI = s.ex_ij[2 * x + 1];
        double em = elevmaskI[iI];
        if (std::isnan(em)) continue;
        double elev = std::max(em, 0.0);
        if (s.interp_style == 0) {
            long k[2]; double w[2];
            linterp_1d_b(g.hcdefs, elev, k, w);
            if (w[0] != 0) out.push_back({x, g.iE(iA, k[0]), s.ex_area[x] * w[0]});
            if (w[1] != 0) out.push_back({x, g.iE(iA, k[1]), s.ex_area[x] * w[1]});
        } else {
            out.push_back({x, g.iE(iA, nearest_1d(g.hcdefs, elev)), s.ex_area[x]});
        }
    }
    return out;
}
static std::vector<STup> gen_GvI(Sheet const &s, const double *elevmaskI)   // L0.cpp:176-189
{
    std::vector<STup> out;
    for (long x = 0; x < s.nX; ++x) {
        long iI = s.ex_ij[2 * x + 1];
        if (!std::isnan(elevmaskI[iI])) out.push_back({x, iI, s.ex_area[x]});
    }
    return out;
}
static std::vector<STup> gen_GvAp(Sheet const &s, const double *elevmaskI)   // L0.cpp:192-214
{
    std::vector<STup> out;
    for (long x = 0; x < s.nX; ++x) {
        long iA = s.ex_ij[2 * x], iI = s.ex_ij[2 * x + 1];
        if (!std::isnan(elevmaskI[iI]) && s.ex_area[x] > 0) out.push_back({x, iA, s.ex_area[x]});
    }
    return out;
}
static std::vector<STup> gen_sApvA(Gcm const &g, Sheet const &s)   // IceRegridder.cpp:41-48
{
    std::vector<STup> out;
    for (size_t id = 0; id < g.A.dim.size(); ++id)
        out.push_back({g.A.dim[id], g.A.dim[id], g.A.native_area[id] / s.gridA_proj_area[id]});
    return out;
}
static std::vector<STup> gen_sEpvE(Gcm const &g, Sheet const &s)   // IceRegridder.cpp:52-65
{
    std::vector<STup> out;
    for (size_t id = 0; id < g.A.dim.size(); ++id)
        for (long k = 0; k < g.nhc(); ++k) {
            long e = g.iE(g.A.dim[id], k);
            out.push_back({e, e, g.A.native_area[id] / s.gridA_proj_area[id]});
        }
    return out;
}

// UrAE (RegridMatrices_Dynamic.cpp:32-46): the A / E symmetry
struct Ur {
    long nfull;
    std::vector<STup> GvAp;    // generator output, already evaluated on this elevmaskI
    std::vector<STup> sApvA;
};

struct Params { bool scale = true; bool correctA = false; double sigma[3] = {0, 0, 0}; bool smooth() const { return sigma[0] != 0; } };

// linear::Weighted_Eigen
struct Weighted {
    Csc M;
    std::vector<double> wM, Mw;
    SparseSet dims[2];
    bool conservative = true;
    bool scaled = true;
};

// smoother.cpp:8-99.  Neighbour search via uniform bins instead of an RTree: the accepted SET
// {j : sum_k ((c_j[k]-c_i[k])/sigma_k)^2 < 4} is identical; only the order of denom_sum differs.
static std::vector<Trip> smoothing_matrix(Sheet const &s, SparseSet const &dimX, const double *elev_s,
    std::vector<double> const &area_d, const double sigma[3])
{
    struct Tup { int d; double c[3]; double area; };
    std::vector<Tup> tup;
    for (size_t id = 0; id < s.dimI.size(); ++id) {
        long iX_s = s.dimI[id];
        if (!dimX.in_sparse(iX_s)) continue;
        double elev = elev_s[iX_s];
        if (std::isnan(elev)) continue;
        int d = dimX.to_dense(iX_s);
        double area = area_d[d];
        if (area == 0.) fail("Area of cell %g must be non-zero\n", (double)iX_s);
        tup.push_back({d, {s.centroid_xy[2 * id], s.centroid_xy[2 * id + 1], elev}, area});
    }
    const double nsigma = 2., nsig2 = nsigma * nsigma;
    std::vector<Trip> out;
    if (tup.empty()) return out;
    // bins of size nsigma*sigma in x and y
    double bx = nsigma * sigma[0], by = nsigma * sigma[1];
    double x0 = tup[0].c[0], y0 = tup[0].c[1], x1 = x0, y1 = y0;
    for (auto &t : tup) { x0 = std::min(x0, t.c[0]); x1 = std::max(x1, t.c[0]); y0 = std::min(y0, t.c[1]); y1 = std::max(y1, t.c[1]); }
    bool fx = std::isfinite(bx) && bx > 0, fy = std::isfinite(by) && by > 0;
    long nbx = fx ? std::min<long>(4096, (long)((x1 - x0) / bx) + 1) : 1;
    long nby = fy ? std::min<long>(4096, (long)((y1 - y0) / by) + 1) : 1;
    auto binx = [&](double x) { return fx ? std::min<long>(nbx - 1, (long)((x - x0) / bx)) : 0; };
    auto biny = [&](double y) { return fy ? std::min<long>(nby - 1, (long)((y - y0) / by)) : 0; };
    std::vector<std::vector<int>> bins(nbx * nby);
    for (size_t t = 0; t < tup.size(); ++t) bins[biny(tup[t].c[1]) * nbx + binx(tup[t].c[0])].push_back((int)t);

    std::vector<std::pair<int, double>> raw;
    for (auto &t0 : tup) {
        raw.clear();
        double denom = 0;
        long cx = binx(t0.c[0]), cy = biny(t0.c[1]);
        for (long yy = std::max<long>(0, cy - 1); yy <= std::min<long>(nby - 1, cy + 1); ++yy)
        for (long xx = std::max<long>(0, cx - 1); xx <= std::min<long>(nbx - 1, cx + 1); ++xx)
        for (int tj : bins[yy * nbx + xx]) {
            Tup const &t = tup[tj];
            double nd2 = 0;
            for (int k = 0; k < 3; ++k) {
                double d = (t.c[k] - t0.c[k]) / sigma[k];
                nd2 += d * d;
            }
            if (nd2 < nsig2) {
                double w = std::exp(-.5 * nd2) * t.area;
                raw.push_back({t.d, w});
                denom += w;
            }
        }
        double factor = 1. / denom;
        for (auto &p : raw) out.push_back({t0.d, p.first, factor * p.second});
    }
    return out;
}

// RegridMatrices_Dynamic.cpp:50-151
static void compute_AEvI(Weighted &ret, Gcm const &g, Sheet const &s, Params const &p,
    const double *elevmaskI, char Igrid, Ur const &AE)
{
    ret.conservative = true;
    SparseSet *dimA = &ret.dims[0], *dimI = &ret.dims[1];
    SparseSet dimG_;
    SparseSet *dimG = &dimG_;
    dimA->sparse_extent = AE.nfull;
    dimI->sparse_extent = (Igrid == 'I' ? s.nI : s.nX);
    dimG->sparse_extent = s.nX;

    SparseSet *d0 = (Igrid == 'I' ? dimG : dimI);
    Csc ApvG = to_eigen(densify(AE.GvAp, ADD_DENSE, d0, dimA), d0, dimA, 'T');
    Csc ApvI;
    if (Igrid == 'I') {
        auto t = densify(gen_GvI(s, elevmaskI), ADD_DENSE, dimG, dimI);
        Csc GvI = to_eigen(t, dimG, dimI, '.');
        auto sGvI = sum(GvI, 0, '-');
        ApvI = mul(scale_cols(ApvG, sGvI), GvI);
    } else {
        ApvI = std::move(ApvG);
    }
    ret.Mw = sum(ApvI, 1, '+');

    if (p.correctA) {
        Csc wAvAp = to_eigen(densify(AE.sApvA, TO_DENSE_IGNORE_MISSING, dimA, dimA), dimA, dimA, '.');
        auto wApvI = sum(ApvI, 0, '+');
        Csc wAvI = scale_cols(wAvAp, wApvI);
        ret.wM = sum(wAvI, 0, '+');
        auto sAvAp = sum(wAvAp, 0, '-');
        if (p.scale) {
            auto sApvI = invert1(wApvI);
            std::vector<double> m(sAvAp.size());
            for (size_t i = 0; i < m.size(); ++i) m[i] = sAvAp[i] * sApvI[i];
            ret.M = scale_rows(m, ApvI);
        } else {
            ret.M = std::move(ApvI);
        }
    } else {
        ret.wM = sum(ApvI, 0, '+');
        if (p.scale) ret.M = scale_rows(sum(ApvI, 0, '-'), ApvI);
        else ret.M = std::move(ApvI);
    }
}

// RegridMatrices_Dynamic.cpp:154-252
static void compute_IvAE(Weighted &ret, Gcm const &g, Sheet const &s, Params const &p,
    const double *elevmaskI, char Igrid, Ur const &AE)
{
    ret.conservative = !p.smooth();
    SparseSet *dimA = &ret.dims[1], *dimI = &ret.dims[0];
    SparseSet dimG_;
    SparseSet *dimG = &dimG_;
    dimA->sparse_extent = AE.nfull;
    dimI->sparse_extent = (Igrid == 'I' ? s.nI : s.nX);
    dimG->sparse_extent = s.nX;

    SparseSet *d0 = (Igrid == 'I' ? dimG : dimI);
    Csc GvAp = to_eigen(densify(AE.GvAp, ADD_DENSE, d0, dimA), d0, dimA, '.');
    Csc IvAp;
    if (Igrid == 'I') {
        Csc IvG = to_eigen(densify(gen_GvI(s, elevmaskI), ADD_DENSE, dimG, dimI), dimG, dimI, 'T');
        auto sGvAp = sum(GvAp, 0, '-');
        IvAp = mul(scale_cols(IvG, sGvAp), GvAp);
    } else {
        IvAp = std::move(GvAp);
    }
    ret.wM = sum(IvAp, 0, '+');

    if (p.correctA) {
        Csc sApvA = to_eigen(densify(AE.sApvA, TO_DENSE_IGNORE_MISSING, dimA, dimA), dimA, dimA, '.');
        auto IvApw = sum(IvAp, 1, '+');
        Csc Aw = scale_cols(sApvA, IvApw);
        ret.Mw = sum(Aw, 0, '+');
        if (p.scale) ret.M = mul(scale_rows(sum(IvAp, 0, '-'), IvAp), sApvA);
        else ret.M = mul(IvAp, sApvA);
    } else {
        ret.Mw = sum(IvAp, 1, '+');
        if (p.scale) ret.M = scale_rows(sum(IvAp, 0, '-'), IvAp);
        else ret.M = std::move(IvAp);
    }

    if (p.smooth()) {
        int n = dimI->dense_extent();
        auto t = smoothing_matrix(s, *dimI, elevmaskI, ret.wM, p.sigma);
        Csc smoothI = assemble(t, n, n);
        ret.M = mul(smoothI, ret.M);
    }
}

// RegridMatrices_Dynamic.cpp:254-332.  (E,A) = (urE,urA) gives EvA; (urA,urE) gives AvE.
static void compute_EvA(Weighted &ret, Gcm const &g, Sheet const &s, Params const &p, Ur const &E, Ur const &A)
{
    ret.conservative = true;
    SparseSet *dimE = &ret.dims[0], *dimA = &ret.dims[1];
    SparseSet dimG_;
    SparseSet *dimG = &dimG_;
    dimA->sparse_extent = A.nfull;
    dimE->sparse_extent = E.nfull;
    dimG->sparse_extent = s.nX;

    auto tA = densify(A.GvAp, ADD_DENSE, dimG, dimA);
    auto tE = densify(E.GvAp, ADD_DENSE, dimG, dimE);
    Csc GvAp = to_eigen(tA, dimG, dimA, '.');
    Csc EpvG = to_eigen(tE, dimG, dimE, 'T');
    auto sGvAp = sum(GvAp, 0, '-');
    Csc EpvAp = mul(scale_cols(EpvG, sGvAp), GvAp);

    auto wEpvAp = sum(EpvAp, 0, '+');
    if (p.correctA) {
        Csc sApvA = to_eigen(densify(A.sApvA, TO_DENSE_IGNORE_MISSING, dimA, dimA), dimA, dimA, '.');
        Csc wEvEp = to_eigen(densify(E.sApvA, TO_DENSE_IGNORE_MISSING, dimE, dimE), dimE, dimE, '.');
        Csc wEvAp = scale_cols(wEvEp, wEpvAp);
        ret.wM = sum(wEvAp, 0, '+');
        auto EpvApw = sum(EpvAp, 1, '+');
        Csc Aw = scale_cols(sApvA, EpvApw);
        ret.Mw = sum(Aw, 0, '+');
        if (p.scale) ret.M = mul(scale_rows(sum(wEvAp, 0, '-'), EpvAp), sApvA);
        else ret.M = mul(EpvAp, sApvA);
    } else {
        ret.wM = wEpvAp;
        ret.Mw = sum(EpvAp, 1, '+');
        if (p.scale) ret.M = scale_rows(invert1(wEpvAp), EpvAp);
        else ret.M = std::move(EpvAp);
    }
}

// GCMRegridder_Standard::regrid_matrices + RegridMatrices_Dynamic::matrix_d
// (RegridMatrices_Dynamic.cpp:334-423).  dims are pre-populated by the caller (or empty = matrix()).
static std::unique_ptr<Weighted> matrix_d(Gcm const &g, int sheet, const double *elevmaskI_in,
    std::string const &spec, Params const &p, SparseSet const &dim0, SparseSet const &dim1)
{
    if (sheet < 0 || sheet >= (int)g.sheets.size()) fail("bad sheet index");
    Sheet const &s = *g.sheets[sheet];
    for (long x = 0; x < s.nX; ++x)
        if (!(s.ex_area[x] > 0)) fail("oracle requires strictly positive exchange-cell areas");
    std::vector<double> elevmaskI(elevmaskI_in, elevmaskI_in + s.nI);   // :352 copies
    std::unique_ptr<Weighted> W(new Weighted);
    W->dims[0] = dim0;
    W->dims[1] = dim1;
    char b = spec.size() == 3 ? spec[0] : 0, a = spec.size() == 3 ? spec[2] : 0;
    if (spec.size() != 3 || spec[1] != 'v') throw std::out_of_range("map::at: " + spec);
    auto urA = [&]() { return Ur{g.A.nA, gen_GvAp(s, elevmaskI.data()), gen_sApvA(g, s)}; };
    auto urE = [&]() { return Ur{g.nE(), gen_GvEp(g, s, elevmaskI.data()), gen_sEpvE(g, s)}; };
    if ((b == 'A' || b == 'E') && (a == 'I' || a == 'X'))
        compute_AEvI(*W, g, s, p, elevmaskI.data(), a, b == 'A' ? urA() : urE());
    else if ((b == 'I' || b == 'X') && (a == 'A' || a == 'E'))
        compute_IvAE(*W, g, s, p, elevmaskI.data(), b, a == 'A' ? urA() : urE());
    else if (b == 'E' && a == 'A') compute_EvA(*W, g, s, p, urE(), urA());
    else if (b == 'A' && a == 'E') compute_EvA(*W, g, s, p, urA(), urE());
    else throw std::out_of_range("map::at: " + spec);
    W->scaled = p.scale;
    return W;
}

// ibmisc linear::Weighted_Eigen::apply (not in tree; SURVEY 8a a16, matrix_formats.rst:150-165):
//   B0 = M * A ; if (!conservative && force_conservation) B[:,n] = B0[:,n] * (Mw.A[:,n])/(wM.B0[:,n])
// X, Y are field-major [nvar][ncell_dense] (Eigen col-major [ncell x nvar], IceCoupler.cpp:231-233).
static void apply(Weighted const &W, const double *X, int nvar, double *Y, bool force_conservation)
{
    Csc const &M = W.M;
    for (int n = 0; n < nvar; ++n) {
        const double *x = X + (size_t)n * M.ncol;
        double *y = Y + (size_t)n * M.nrow;
        for (int i = 0; i < M.nrow; ++i) y[i] = 0;
        for (int j = 0; j < M.ncol; ++j) {
            double xj = x[j];
            for (int q = M.ptr[j]; q < M.ptr[j + 1]; ++q) y[M.row[q]] += M.val[q] * xj;
        }
        if (W.conservative || !force_conservation) continue;
        double TA = 0, TB = 0;
        for (int j = 0; j < M.ncol; ++j) TA += W.Mw[j] * x[j];
        for (int i = 0; i < M.nrow; ++i) TB += W.wM[i] * y[i];
        double factor = TA * (1. / TB);
        for (int i = 0; i < M.nrow; ++i) y[i] = y[i] * factor;
    }
}

// ---------------------------------------------------------------------------------------
// Hntr (modele/hntr.cpp:63-168, hntr.hpp:204-435).  Arrays keep the reference's 1-based VALUES
// (IMIN.. hold 1-based A indices) but are stored 0-based: X(IB) -> x[IB-1]; SINA(JA) -> sina[JA].
struct Hntr {
    int imA, jmA, imB, jmB;
    double offiA, dlatA, offiB, dlatB, datmis;
    std::vector<double> sina, sinb, fmin, fmax, gmin, gmax;
    std::vector<int> imin, imax, jmin, jmax;
    std::vector<double> dxypA, dxypB;
};

static std::vector<double> make_dxyp(int im, int jm)   // hntr.cpp:34-52
{
    std::vector<double> d(jm);
    double dLON = (2. * M_PI) / im, dLAT = M_PI / jm;
    for (int j = 1; j <= jm; ++j) {
        double SINS = sin(dLAT * (j - jm / 2 - 1));
        double SINN = sin(dLAT * (j - jm / 2));
        d[j - 1] = dLON * (SINN - SINS);
    }
    return d;
}

static void hntr_spec_check(int im, int jm, double offi, double dlat)   // GridSpec.cpp:125-132
{
    if (dlat < .1 * (180. * 60. / jm)) fail("dlat in HntrGrid seems to small; it should be in MINUTES ,not DEGREES: %g", .1 * (180. * 60. / jm));
}

static std::unique_ptr<Hntr> hntr_new(int imB, int jmB, double offiB, double dlatB,
    int imA, int jmA, double offiA, double dlatA, double datmis)
{
    hntr_spec_check(imB, jmB, offiB, dlatB);
    hntr_spec_check(imA, jmA, offiA, dlatA);
    std::unique_ptr<Hntr> h(new Hntr);
    h->imA = imA; h->jmA = jmA; h->imB = imB; h->jmB = jmB;
    h->offiA = offiA; h->dlatA = dlatA; h->offiB = offiB; h->dlatB = dlatB; h->datmis = datmis;
    h->sina.assign(jmA + 1, 0); h->sinb.assign(jmB + 1, 0);
    h->fmin.assign(imB, 0); h->fmax.assign(imB, 0); h->imin.assign(imB, 0); h->imax.assign(imB, 0);
    h->gmin.assign(jmB, 0); h->gmax.assign(jmB, 0); h->jmin.assign(jmB, 0); h->jmax.assign(jmB, 0);
    h->dxypA = make_dxyp(imA, jmA);
    h->dxypB = make_dxyp(imB, jmB);

    // east-west (hntr.cpp:84-115): domain scaled to [0, imA*imB)
    {
        double width = imB;                    // one A cell in the scaled domain
        int ia = 1;
        double edgeA = (ia + offiA - imA) * imB;   // eastern edge of A cell ia
        int ibPrev = imB;
        for (int ib = 1; ib <= imB; ++ib) {
            double edgeB = (ib - 1 + offiB) * imA; // eastern edge of B cell ib-1 (= western edge of ib)
            while (edgeA < edgeB) { ia += 1; edgeA += width; }
            if (edgeA == edgeB) {
                h->imax[ibPrev - 1] = ia; h->fmax[ibPrev - 1] = 0;
                ia += 1; edgeA += width;
                h->imin[ib - 1] = ia; h->fmin[ib - 1] = 0;
            } else {
                h->imax[ibPrev - 1] = ia; h->fmax[ibPrev - 1] = (edgeA - edgeB) / width;
                h->imin[ib - 1] = ia; h->fmin[ib - 1] = 1 - h->fmax[ibPrev - 1];
            }
            ibPrev = ib;
        }
        h->imax[imB - 1] += imA;
    }
    // north-south (hntr.cpp:118-168): sines of northern cell edges; dlat in minutes
    {
        const double MIN_TO_RAD = (2. * M_PI) / (360 * 60);
        double eqA = .5 * (1 + jmA);
        for (int ja = 1; ja <= jmA - 1; ++ja) h->sina[ja] = sin(((ja + .5 - eqA) * dlatA) * MIN_TO_RAD);
        h->sina[0] = -1; h->sina[jmA] = 1;
        double eqB = .5 * (1 + jmB);
        for (int jb = 1; jb <= jmB - 1; ++jb) h->sinb[jb] = sin(((jb + .5 - eqB) * dlatB) * MIN_TO_RAD);
        h->sinb[0] = -1; h->sinb[jmB] = 1;

        h->jmin[0] = 1; h->gmin[0] = 0;
        int ja = 1;
        for (int jb = 1; jb <= jmB - 1; ++jb) {
            while (h->sina[ja] < h->sinb[jb]) ++ja;
            if (h->sina[ja] == h->sinb[jb]) {
                h->jmax[jb - 1] = ja; h->gmax[jb - 1] = 0;
                ja += 1;
                h->jmin[jb] = ja; h->gmin[jb] = 0;
            } else {
                h->jmax[jb - 1] = ja; h->gmax[jb - 1] = h->sina[ja] - h->sinb[jb];
                h->jmin[jb] = ja; h->gmin[jb] = h->sinb[jb] - h->sina[ja - 1];
            }
        }
        h->jmax[jmB - 1] = jmA; h->gmax[jmB - 1] = 0;
    }
    return h;
}

// The engine (hntr.hpp:204-244): visit(IJB0, JB, IJA0, FG) for every overlapping pair then done(IJB0, JB).
template <class OnA, class OnB>
static void hntr_walk(Hntr const &h, OnA onA, OnB onB)
{
    for (int JB = 1; JB <= h.jmB; ++JB) {
        int jlo = h.jmin[JB - 1], jhi = h.jmax[JB - 1];
        for (int IB = 1; IB <= h.imB; ++IB) {
            long ijb0 = (IB - 1) + (long)h.imB * (JB - 1);
            int ilo = h.imin[IB - 1], ihi = h.imax[IB - 1];
            for (int JA = jlo; JA <= jhi; ++JA) {
                double G = h.sina[JA] - h.sina[JA - 1];
                if (JA == jlo) G -= h.gmin[JB - 1];
                if (JA == jhi) G -= h.gmax[JB - 1];
                for (int ir = ilo; ir <= ihi; ++ir) {
                    int IA = 1 + ((ir - 1) % h.imA);
                    long ija0 = (IA - 1) + (long)h.imA * (JA - 1);
                    double F = 1;
                    if (ir == ilo) F -= h.fmin[IB - 1];
                    if (ir == ihi) F -= h.fmax[IB - 1];
                    onA(ija0, F * G);
                }
            }
            onB(ijb0, JB);
        }
    }
}

// Hntr::regrid (hntr.hpp:385-435) with RegridAccum (:342-380).  wta == nullptr means a constant
// weight array of 1 (etopo1_ice.cpp:24-26 const_array).
template <class SrcT>
static void hntr_regrid(Hntr const &h, const double *wta, const SrcT *A, double *B, bool mean_polar, double wtm, double wtb)
{
    double WEIGHT = 0, VALUE = 0;
    hntr_walk(h,
        [&](long ija, double FG) {
            double w = wtm * (wta ? wta[ija] : 1.0) + wtb;
            double wt = FG * w;
            WEIGHT += wt;
            VALUE += wt * (double)A[ija];
        },
        [&](long ijb, int) {
            B[ijb] = (WEIGHT == 0 ? h.datmis : VALUE / WEIGHT);
            WEIGHT = 0; VALUE = 0;
        });
    if (mean_polar) {
        for (int JB = 1; JB <= h.jmB; JB += std::max(1, h.jmB - 1)) {
            double mean = h.datmis, cnt = 0, tot = 0;
            double *row = B + (long)h.imB * (JB - 1);
            int IB = 0;
            for (; IB < h.imB; ++IB) {
                if (row[IB] == h.datmis) break;     // any missing cell => whole row missing (hntr.hpp:419-427)
                cnt += 1; tot += row[IB];
            }
            if (IB == h.imB && cnt != 0) mean = tot / cnt;
            for (IB = 0; IB < h.imB; ++IB) row[IB] = mean;
            if (h.jmB == 1) break;
        }
    }
}

// Hntr::overlap (hntr.hpp:249-293) / scaled_regrid_matrix (:295-339): R2 < 0 selects the scaled flavour.
static void hntr_matrix(Hntr const &h, double R2, std::vector<long> &rows, std::vector<long> &cols, std::vector<double> &vals)
{
    std::vector<std::pair<long, double>> buf;
    double WEIGHT = 0;
    hntr_walk(h,
        [&](long ija, double FG) { WEIGHT += FG; buf.push_back({ija, FG}); },
        [&](long ijb, int JB) {
            double by = 1. / WEIGHT;
            for (auto &e : buf) {
                double v;
                if (R2 < 0) v = e.second * by;
                else { double area = R2 * h.dxypB[JB - 1]; v = e.second * by * area; }
                rows.push_back(ijb); cols.push_back(e.first); vals.push_back(v);
            }
            buf.clear(); WEIGHT = 0;
        });
}

// ---------------------------------------------------------------------------------------
// Coupling step, regridding side.
//
// Variable transformation in front of a regrid (IceCoupler.cpp:203-252, :404-446):
//   Z = O * M + b.replicate(n,1)   with O dense [ncell x nin], M sparse col-major [nin x nout], b [nout].
// Eigen's dense*sparse walks the stored entries of each column of M in order; structural zeros are never touched.
static void vartrans(const double *O, long ncell, int nin, int nout, const int *colptr, const int *row, const double *val,
                     const double *b, double *Z)
{
    for (int k = 0; k < nout; ++k)
        for (int j = colptr[k]; j < colptr[k + 1]; ++j)
            if (row[j] < 0 || row[j] >= nin) fail("vartrans: row index out of range");
    for (long i = 0; i < ncell; ++i)
        for (int k = 0; k < nout; ++k) {
            double acc = 0.0;
            for (int j = colptr[k]; j < colptr[k + 1]; ++j) acc += O[i * nin + row[j]] * val[j];
            Z[i * nout + k] = acc + b[k];
        }
}

static Csc transpose(Csc const &M)
{
    Csc T;
    T.nrow = M.ncol; T.ncol = M.nrow;
    T.ptr.assign(M.nrow + 1, 0);
    for (int r : M.row) T.ptr[r + 1]++;
    for (int i = 0; i < M.nrow; ++i) T.ptr[i + 1] += T.ptr[i];
    T.row.resize(M.row.size()); T.val.resize(M.val.size());
    std::vector<int> fill(T.ptr.begin(), T.ptr.end() - 1);
    for (int j = 0; j < M.ncol; ++j)
        for (int p = M.ptr[j]; p < M.ptr[j + 1]; ++p) { int q = fill[M.row[p]]++; T.row[q] = j; T.val[q] = M.val[p]; }
    return T;
}
// Eigen sparse - sparse (CwiseBinaryOp<scalar_difference_op>): union of the patterns; a - 0 and 0 - b where one is missing.
static Csc sub(Csc const &A, Csc const &B)
{
    if (A.nrow != B.nrow || A.ncol != B.ncol) fail("sub: shapes differ");
    Csc C;
    C.nrow = A.nrow; C.ncol = A.ncol;
    C.ptr.assign(A.ncol + 1, 0);
    for (int j = 0; j < A.ncol; ++j) {
        int p = A.ptr[j], q = B.ptr[j];
        const int pe = A.ptr[j + 1], qe = B.ptr[j + 1];
        while (p < pe || q < qe) {
            if (q >= qe || (p < pe && A.row[p] < B.row[q])) { C.row.push_back(A.row[p]); C.val.push_back(A.val[p] - 0.); ++p; }
            else if (p >= pe || B.row[q] < A.row[p]) { C.row.push_back(B.row[q]); C.val.push_back(0. - B.val[q]); ++q; }
            else { C.row.push_back(A.row[p]); C.val.push_back(A.val[p] - B.val[q]); ++p; ++q; }
        }
        C.ptr[j + 1] = (int)C.row.size();
    }
    return C;
}
// GCMCoupler.cpp:388-449 sparsify(XuE, {ID, TO_SPARSE}, {nX, -1}): rows keep their (identity) ids, columns and Mw go
// to sparse E indexing.
struct SparsifiedX { Csc M; std::vector<double> wM, Mw; };
static SparsifiedX sparsify_XuE(Weighted const &W, long nE)
{
    SparsifiedX S;
    std::vector<Trip> t;
    for (int j = 0; j < W.M.ncol; ++j)
        for (int p = W.M.ptr[j]; p < W.M.ptr[j + 1]; ++p) t.push_back(Trip{W.M.row[p], (int)W.dims[1].to_s[j], W.M.val[p]});
    S.M = assemble(t, W.M.nrow, (int)nE);
    S.wM = W.wM;
    S.Mw.assign(nE, 0.0);
    for (int j = 0; j < W.M.ncol; ++j) S.Mw[W.dims[1].to_s[j]] = W.Mw[j];
    return S;
}
// e1ve0.cpp:55-106 compute_E1vE0c + :32-52 consolidate.  Entries come out sorted by (iE1, iE0); explicit zeros are kept
// (whether ibmisc's spcopy drops them is not visible in tree: unpinned, and harmless -- they are zeros).
static void compute_E1vE0c(std::vector<Weighted const *> const &XuE1s, std::vector<Weighted const *> const &XuE0s, long nE,
                           std::vector<long> &rows, std::vector<long> &cols, std::vector<double> &vals)
{
    struct T3 { long r, c; double v; };
    std::vector<T3> E1vE0c;
    std::vector<double> sE1(nE, 0.0);
    for (size_t i = 0; i < XuE1s.size(); ++i) {
        if (XuE1s[i]->M.nrow != XuE0s[i]->M.nrow) fail("compute_E1vE0c: exchange grids differ");
        SparsifiedX X1 = sparsify_XuE(*XuE1s[i], nE), X0 = sparsify_XuE(*XuE0s[i], nE);
        std::vector<double> sXuE0 = invert1(X0.wM), sXuE1 = invert1(X1.wM);
        Csc E1uX = transpose(X1.M);
        Csc XvE1 = scale_rows(sXuE1, X1.M);
        Csc XvE0 = scale_rows(sXuE0, X0.M);
        Csc local = mul(E1uX, sub(XvE0, XvE1));
        for (int j = 0; j < local.ncol; ++j)
            for (int p = local.ptr[j]; p < local.ptr[j + 1]; ++p) E1vE0c.push_back(T3{local.row[p], j, local.val[p]});
        for (long e = 0; e < nE; ++e) sE1[e] += X1.Mw[e];
    }
    for (double &x : sE1) x = 1. / x;
    for (T3 &t : E1vE0c) t.v *= sE1[t.r];
    std::stable_sort(E1vE0c.begin(), E1vE0c.end(), [](T3 const &a, T3 const &b) { return a.r != b.r ? a.r < b.r : a.c < b.c; });
    rows.clear(); cols.clear(); vals.clear();
    for (size_t i = 0; i < E1vE0c.size(); ++i) {
        if (!rows.empty() && rows.back() == E1vE0c[i].r && cols.back() == E1vE0c[i].c) vals.back() += E1vE0c[i].v;
        else { rows.push_back(E1vE0c[i].r); cols.push_back(E1vE0c[i].c); vals.push_back(E1vE0c[i].v); }
    }
}

// ---------------------------------------------------------------------------------------
// Exchange grid of projected GCM polygons with a rectilinear ice grid.
// gridgen/GridGen_Exchange.cpp:175-316: for every GCM cell, the ice cells under its bounding box (RTree search), the
// overlap polygon (CGAL there; here a rectangle window, so Sutherland-Hodgman clipping against its four sides is exact
// up to rounding) and its area by the shoelace formula (Cell::proj_area).  CGAL is not in this image: parity unpinned.
typedef std::vector<std::pair<double, double>> Poly;
static Poly clip_halfplane(Poly const &in, int axis, bool keep_ge, double c)
{
    Poly out;
    for (size_t i = 0; i < in.size(); ++i) {
        auto const &p = in[i];
        auto const &q = in[(i + 1) % in.size()];
        const double pv = axis == 0 ? p.first : p.second, qv = axis == 0 ? q.first : q.second;
        const bool pin = keep_ge ? pv >= c : pv <= c, qin = keep_ge ? qv >= c : qv <= c;
        if (pin) out.push_back(p);
        if (pin != qin) {
            const double t = (c - pv) / (qv - pv);
            if (axis == 0) out.push_back({c, p.second + t * (q.second - p.second)});
            else out.push_back({p.first + t * (q.first - p.first), c});
        }
    }
    return out;
}
static double poly_area(Poly const &v)
{
    double s = 0.0;
    for (size_t i = 0; i < v.size(); ++i) {
        auto const &p = v[i];
        auto const &q = v[(i + 1) % v.size()];
        s += p.first * q.second - q.first * p.second;
    }
    s *= 0.5;
    return s < 0 ? -s : s;
}
static void exgrid_rect(int npoly, const int *pptr, const double *pxy, const long *iA, int nx, const double *xe, int ny,
                        const double *ye, bool major_x, std::vector<int> &ij, std::vector<double> &area)
{
    ij.clear(); area.clear();
    for (int p = 0; p < npoly; ++p) {
        Poly poly;
        for (int v = pptr[p]; v < pptr[p + 1]; ++v) poly.push_back({pxy[2 * v], pxy[2 * v + 1]});
        double x0 = poly[0].first, x1 = x0, y0 = poly[0].second, y1 = y0;
        for (auto const &v : poly) { x0 = std::min(x0, v.first); x1 = std::max(x1, v.first); y0 = std::min(y0, v.second); y1 = std::max(y1, v.second); }
        std::vector<std::pair<long, double>> cells;      // (iI, area)
        for (int ix = 0; ix < nx; ++ix) {
            if (xe[ix + 1] <= x0 || xe[ix] >= x1) continue;
            for (int iy = 0; iy < ny; ++iy) {
                if (ye[iy + 1] <= y0 || ye[iy] >= y1) continue;
                Poly c = clip_halfplane(poly, 0, true, xe[ix]);
                if (c.size() >= 3) c = clip_halfplane(c, 0, false, xe[ix + 1]);
                if (c.size() >= 3) c = clip_halfplane(c, 1, true, ye[iy]);
                if (c.size() >= 3) c = clip_halfplane(c, 1, false, ye[iy + 1]);
                if (c.size() < 3) continue;
                const double a = poly_area(c);
                if (a > 0) cells.push_back({major_x ? (long)ix * ny + iy : (long)iy * nx + ix, a});
            }
        }
        std::sort(cells.begin(), cells.end());
        for (auto const &c : cells) { ij.push_back((int)iA[p]); ij.push_back((int)c.first); area.push_back(c.second); }
    }
}

}   // namespace orc

// =======================================================================================
// C ABI (ctypes).  Every call returns 0 on success, -1 on error (message via orc_last_error()).
using namespace orc;
static thread_local std::string g_err;
#define ORC_TRY try {
#define ORC_CATCH(rv) } catch (std::exception const &e) { g_err = e.what(); return rv; }

extern "C" {

const char *orc_last_error() { return g_err.c_str(); }

void *orc_gcm_new(long nA, int nreal, const long *dimA, const double *native_area, int nhc, const double *hcdefs)
{
    ORC_TRY
    Gcm *g = new Gcm;
    g->A.nA = nA;
    g->A.dim.assign(dimA, dimA + nreal);
    g->A.native_area.assign(native_area, native_area + nreal);
    g->hcdefs.assign(hcdefs, hcdefs + nhc);
    return g;
    ORC_CATCH(nullptr)
}
void orc_gcm_free(void *g) { delete (Gcm *)g; }

int orc_add_sheet(void *g_, long nI, int nreal_I, const long *dimI, const double *native_areaI, const double *centroid_xy,
    long nX, const int *ex_ij, const double *ex_area, const double *gridA_proj_area, int interp_style)
{
    ORC_TRY
    Gcm *g = (Gcm *)g_;
    std::unique_ptr<Sheet> s(new Sheet);
    s->nI = nI;
    s->dimI.assign(dimI, dimI + nreal_I);
    s->native_areaI.assign(native_areaI, native_areaI + nreal_I);
    if (centroid_xy) s->centroid_xy.assign(centroid_xy, centroid_xy + 2 * (size_t)nreal_I);
    s->nX = nX;
    s->ex_ij.assign(ex_ij, ex_ij + 2 * nX);
    s->ex_area.assign(ex_area, ex_area + nX);
    s->gridA_proj_area.assign(gridA_proj_area, gridA_proj_area + g->A.dim.size());
    s->interp_style = interp_style;
    g->sheets.push_back(std::move(s));
    return (int)g->sheets.size() - 1;
    ORC_CATCH(-1)
}

// n0/n1 < 0: fresh dims (RegridMatrices_Dynamic::matrix); else dims pre-populated with ids0/ids1 (matrix_d).
void *orc_matrix(void *g_, int sheet, const double *elevmaskI, const char *spec, int scale, int correctA,
    const double *sigma, long n0, const long *ids0, long n1, const long *ids1)
{
    ORC_TRY
    Gcm *g = (Gcm *)g_;
    Params p;
    p.scale = scale; p.correctA = correctA;
    if (sigma) for (int k = 0; k < 3; ++k) p.sigma[k] = sigma[k];
    SparseSet d0, d1;
    for (long i = 0; i < n0; ++i) d0.add_dense(ids0[i]);
    for (long i = 0; i < n1; ++i) d1.add_dense(ids1[i]);
    return matrix_d(*g, sheet, elevmaskI, spec, p, d0, d1).release();
    ORC_CATCH(nullptr)
}
void orc_w_free(void *w) { delete (Weighted *)w; }
// info[0..5] = nrow, ncol, nnz, conservative, sparse_extent0, sparse_extent1
void orc_w_info(void *w_, long *info)
{
    Weighted *w = (Weighted *)w_;
    info[0] = w->M.nrow; info[1] = w->M.ncol; info[2] = w->M.nnz(); info[3] = w->conservative;
    info[4] = w->dims[0].sparse_extent; info[5] = w->dims[1].sparse_extent;
}
// COO in CSC storage order (columns ascending, rows ascending inside)
void orc_w_coo(void *w_, int *row, int *col, double *val)
{
    Weighted *w = (Weighted *)w_;
    for (int j = 0; j < w->M.ncol; ++j)
        for (int p = w->M.ptr[j]; p < w->M.ptr[j + 1]; ++p) { row[p] = w->M.row[p]; col[p] = j; val[p] = w->M.val[p]; }
}
void orc_w_weights(void *w_, double *wM, double *Mw)
{
    Weighted *w = (Weighted *)w_;
    std::copy(w->wM.begin(), w->wM.end(), wM);
    std::copy(w->Mw.begin(), w->Mw.end(), Mw);
}
void orc_w_dim(void *w_, int which, long *ids)
{
    Weighted *w = (Weighted *)w_;
    std::copy(w->dims[which].to_s.begin(), w->dims[which].to_s.end(), ids);
}
// X: [nvar][ncol]  Y: [nvar][nrow]  (dense indexing)
int orc_w_apply(void *w_, const double *X, int nvar, double *Y, int force_conservation)
{
    ORC_TRY
    apply(*(Weighted *)w_, X, nvar, Y, force_conservation);
    return 0;
    ORC_CATCH(-1)
}

// VectorMultivec::to_dense_scale + to_dense (multivec.cpp:35-81): merge of several sheets'
// sparse contributions.  index[n], weights[n], vals[n][nvar]  ->  dense[nvar][nE]
int orc_multivec_to_dense(long n, const long *index, const double *weights, const double *vals, int nvar,
    long nE, double fill, double *scaleE, double *dense)
{
    ORC_TRY
    for (long i = 0; i < nE; ++i) scaleE[i] = 0;
    for (long i = 0; i < n; ++i) {
        if (index[i] >= nE) fail("Index out of range: %g vs. %g", (double)index[i], (double)nE);
        scaleE[index[i]] += weights[i];
    }
    for (long i = 0; i < nE; ++i) scaleE[i] = 1. / scaleE[i];
    for (int v = 0; v < nvar; ++v) {
        double *d = dense + (size_t)v * nE;
        for (long i = 0; i < nE; ++i) d[i] = NaN;
        for (long i = 0; i < n; ++i) {
            long e = index[i];
            double t = vals[(size_t)i * nvar + v] * scaleE[e];
            if (std::isnan(d[e])) d[e] = t; else d[e] += t;
        }
        for (long i = 0; i < nE; ++i) if (std::isnan(d[i])) d[i] = fill;
    }
    return 0;
    ORC_CATCH(-1)
}

int orc_vartrans(const double *O, long ncell, int nin, int nout, const int *colptr, const int *row, const double *val,
                 const double *b, double *Z)
{
    ORC_TRY
    vartrans(O, ncell, nin, nout, colptr, row, val, b, Z);
    return 0;
    ORC_CATCH(-1)
}
// Two calls: rows == NULL returns the entry count; then with buffers.
long orc_e1ve0c(int nsheet, void **XuE1, void **XuE0, long nE, long *rows, long *cols, double *vals)
{
    ORC_TRY
    std::vector<Weighted const *> a, b;
    for (int i = 0; i < nsheet; ++i) { a.push_back((Weighted *)XuE1[i]); b.push_back((Weighted *)XuE0[i]); }
    std::vector<long> r, c;
    std::vector<double> v;
    compute_E1vE0c(a, b, nE, r, c, v);
    if (rows) {
        std::copy(r.begin(), r.end(), rows);
        std::copy(c.begin(), c.end(), cols);
        std::copy(v.begin(), v.end(), vals);
    }
    return (long)r.size();
    ORC_CATCH(-1)
}

// Two calls: ex_ij == NULL returns the count.
long orc_exgrid(int npoly, const int *pptr, const double *pxy, const long *iA, int nx, const double *xe, int ny, const double *ye,
                int major_x, int *ex_ij, double *ex_area)
{
    ORC_TRY
    std::vector<int> ij;
    std::vector<double> area;
    exgrid_rect(npoly, pptr, pxy, iA, nx, xe, ny, ye, major_x != 0, ij, area);
    if (ex_ij) { std::copy(ij.begin(), ij.end(), ex_ij); std::copy(area.begin(), area.end(), ex_area); }
    return (long)area.size();
    ORC_CATCH(-1)
}

// GCMRegridder_Standard::filter_cellsA(keepA) (GCMRegridder.cpp:154-164) with the documented intent of
// IceRegridder::filter_cellsA / ExchangeGrid::filter_cellsB (IceRegridder.cpp:150-178, AbbrGrid.cpp:33-49): exchange
// cells of dropped GCM cells go (order kept), then agridA->filter_cells(keepA).  keep[nA], nonzero = keep.
int orc_filter_cellsA(void *g_, const unsigned char *keep)
{
    ORC_TRY
    Gcm *g = (Gcm *)g_;
    for (auto &sp : g->sheets) {
        Sheet &s = *sp;
        std::vector<int> ij;
        std::vector<double> ar;
        for (long x = 0; x < s.nX; ++x)
            if (keep[s.ex_ij[2 * x]]) { ij.push_back(s.ex_ij[2 * x]); ij.push_back(s.ex_ij[2 * x + 1]); ar.push_back(s.ex_area[x]); }
        s.ex_ij.swap(ij); s.ex_area.swap(ar);
        s.nX = (long)s.ex_area.size();
        std::vector<double> proj;
        for (size_t d = 0; d < g->A.dim.size(); ++d) if (keep[g->A.dim[d]]) proj.push_back(s.gridA_proj_area[d]);
        s.gridA_proj_area.swap(proj);
    }
    std::vector<long> dim;
    std::vector<double> area;
    for (size_t d = 0; d < g->A.dim.size(); ++d) if (keep[g->A.dim[d]]) { dim.push_back(g->A.dim[d]); area.push_back(g->A.native_area[d]); }
    g->A.dim.swap(dim); g->A.native_area.swap(area);
    return 0;
    ORC_CATCH(-1)
}
long orc_nX(void *g_, int sheet) { return ((Gcm *)g_)->sheets[sheet]->nX; }

// combine_chunks (modele/combine_global_ec.cpp:95-262): per-chunk unscaled matrices -> one matrix in sparse indexing.
// Buffers: M_* [sum nnz], wM_* [sum nrow], Mw_* [sum ncol].  Entries of a chunk in its COO (column-major) order.
int orc_combine_chunks(int nchunk, void **W_, int scale, long *M_row, long *M_col, double *M_val, long *wM_index, double *wM_val,
                       long *Mw_index, double *Mw_val)
{
    ORC_TRY
    std::vector<Weighted *> W;
    for (int i = 0; i < nchunk; ++i) W.push_back((Weighted *)W_[i]);
    const long shape0 = W[0]->dims[0].sparse_extent;
    std::vector<double> sM_s(shape0, 0.0);
    for (Weighted *w : W)
        for (int i = 0; i < w->M.nrow; ++i) sM_s[w->dims[0].to_s[i]] += w->wM[i];
    for (double &x : sM_s) x = 1. / x;
    long on = 0, orow = 0, ocol = 0;
    for (Weighted *w : W) {
        for (int i = 0; i < w->M.nrow; ++i) { wM_index[orow] = w->dims[0].to_s[i]; wM_val[orow] = w->wM[i]; ++orow; }
        for (int j = 0; j < w->M.ncol; ++j)
            for (int p = w->M.ptr[j]; p < w->M.ptr[j + 1]; ++p) {
                const long iB_s = w->dims[0].to_s[w->M.row[p]], iA_s = w->dims[1].to_s[j];
                M_row[on] = iB_s; M_col[on] = iA_s;
                M_val[on] = scale ? w->M.val[p] * sM_s[iB_s] : w->M.val[p];
                ++on;
            }
        for (int j = 0; j < w->M.ncol; ++j) { Mw_index[ocol] = w->dims[1].to_s[j]; Mw_val[ocol] = w->Mw[j]; ++ocol; }
    }
    return 0;
    ORC_CATCH(-1)
}

int orc_linterp(const double *xp, int n, double xx, long *idx, double *w)
{
    ORC_TRY
    std::vector<double> v(xp, xp + n);
    linterp_1d_b(v, xx, idx, w);
    return 0;
    ORC_CATCH(-1)
}
int orc_nearest(const double *xp, int n, double xx) { std::vector<double> v(xp, xp + n); return nearest_1d(v, xx); }

void *orc_hntr_new(int imB, int jmB, double offiB, double dlatB, int imA, int jmA, double offiA, double dlatA, double datmis)
{
    ORC_TRY
    return hntr_new(imB, jmB, offiB, dlatB, imA, jmA, offiA, dlatA, datmis).release();
    ORC_CATCH(nullptr)
}
void orc_hntr_free(void *h) { delete (Hntr *)h; }
// which: 0 SINA 1 SINB 2 FMIN 3 FMAX 4 GMIN 5 GMAX (double) ; 6 IMIN 7 IMAX 8 JMIN 9 JMAX (int, 1-based values)
// 10 dxypA 11 dxypB
void orc_hntr_get(void *h_, int which, double *dout, int *iout)
{
    Hntr *h = (Hntr *)h_;
    std::vector<double> const *d[] = {&h->sina, &h->sinb, &h->fmin, &h->fmax, &h->gmin, &h->gmax};
    std::vector<int> const *i[] = {&h->imin, &h->imax, &h->jmin, &h->jmax};
    if (which < 6) std::copy(d[which]->begin(), d[which]->end(), dout);
    else if (which < 10) std::copy(i[which - 6]->begin(), i[which - 6]->end(), iout);
    else if (which == 10) std::copy(h->dxypA.begin(), h->dxypA.end(), dout);
    else std::copy(h->dxypB.begin(), h->dxypB.end(), dout);
}
// src_dtype: 0 = f64, 1 = i16.  wta may be NULL (constant 1).
int orc_hntr_regrid(void *h_, const double *wta, const void *A, int src_dtype, double *B, int mean_polar, double wtm, double wtb)
{
    ORC_TRY
    Hntr *h = (Hntr *)h_;
    if (src_dtype == 0) hntr_regrid(*h, wta, (const double *)A, B, mean_polar, wtm, wtb);
    else hntr_regrid(*h, wta, (const int16_t *)A, B, mean_polar, wtm, wtb);
    return 0;
    ORC_CATCH(-1)
}
// R2 < 0: scaled_regrid_matrix; else overlap with R^2.  Two-call protocol: rows==NULL returns the count.
long orc_hntr_matrix(void *h_, double R2, long *rows, long *cols, double *vals)
{
    ORC_TRY
    std::vector<long> r, c; std::vector<double> v;
    hntr_matrix(*(Hntr *)h_, R2, r, c, v);
    if (rows) { std::copy(r.begin(), r.end(), rows); std::copy(c.begin(), c.end(), cols); std::copy(v.begin(), v.end(), vals); }
    return (long)r.size();
    ORC_CATCH(-1)
}

}   // extern "C"
