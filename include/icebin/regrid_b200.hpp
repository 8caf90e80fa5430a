// regrid_b200.hpp -- host-side C++ mirror of the reference's regridding API over the C ABI (icebin_b200.h).
//
// Same class names, argument meaning and error behaviour as the reference (citibeth/icebin, slib/icebin/):
//   icebin::RegridParams              RegridMatrices.hpp:17-36
//   icebin::SparseSetT                spsparse::SparseSet<long,int> (eigen_types.hpp:24)
//   icebin::linear::Weighted          ibmisc::linear::Weighted_Eigen (RegridMatrices_Dynamic.cpp:63; matrix_formats.rst)
//   icebin::RegridMatrices_Dynamic    RegridMatrices_Dynamic.hpp:20-61  (matrix, matrix_d)
//   icebin::IceRegridder              IceRegridder.hpp:60-133 (data holder handed to add_sheet)
//   icebin::GCMRegridder_Standard     GCMRegridder.hpp:312-400 (init, add_sheet, nA/nE/nhc/nI, wA, regrid_matrices)
//   icebin::modele::HntrSpec / Hntr   GridSpec.hpp:143-170, modele/hntr.hpp:63-184
// Differences: Blitz/Eigen containers are replaced by std::vector / raw pointers; grids are given as arrays, not
// NetCDF files; errors throw icebin::Exception (the reference's icebin_error handler throws too, error.hpp:22-32).
// Header-only; link with -licebin_b200.
#pragma once
#include <array>
#include <cmath>
#include <cstdint>
#include <functional>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../icebin_b200.h"

namespace icebin {

struct Exception : std::runtime_error {
    int code;
    Exception(int c, std::string const &m) : std::runtime_error(m), code(c) {}
};
inline void ib_check(long status)
{
    if (status < 0) throw Exception((int)status, ib_last_error());
}

/** Parameters controlling the generation of regridding matrices (RegridMatrices.hpp:17-36) */
struct RegridParams {
    bool scale;
    bool correctA;
    std::array<double, 3> sigma;
    bool smooth() const { return sigma[0] != 0; }
    RegridParams() : scale(true), correctA(false), sigma({0., 0., 0.}) {}
    RegridParams(bool _scale, bool _correctA, std::array<double, 3> const &_sigma = {0., 0., 0.})
        : scale(_scale), correctA(_correctA), sigma(_sigma) {}
};

/** Ordered sparse<->dense index bijection, resident on the GPU. */
class SparseSetT {
    ib_dim *h = nullptr;
    friend class RegridMatrices_Dynamic;
public:
    SparseSetT() { ib_check(ib_dim_create(&h)); }
    SparseSetT(long sparse_extent, std::vector<int64_t> const &ids) { ib_check(ib_dim_create_ids(&h, sparse_extent, (int64_t)ids.size(), ids.data())); }
    /** id_sparse_set<SparseSetT>(n) (IceCoupler.cpp:366) */
    static SparseSetT identity(long n) { SparseSetT s(nullptr); ib_check(ib_dim_create_identity(&s.h, n)); return s; }
    SparseSetT(SparseSetT &&o) noexcept : h(o.h) { o.h = nullptr; }
    SparseSetT(SparseSetT const &) = delete;
    ~SparseSetT() { if (h) ib_dim_free(h); }
    long dense_extent() const { return (long)ib_dim_dense_extent(h); }
    long sparse_extent() const { return (long)ib_dim_sparse_extent(h); }
    std::vector<int64_t> to_sparse() const { std::vector<int64_t> v((size_t)dense_extent()); ib_check(ib_dim_get_ids(h, v.data())); return v; }
private:
    explicit SparseSetT(std::nullptr_t) {}
};

namespace linear {
/** Regrid matrix M (dense-indexed CSR) + weight vectors wM (rows) and Mw (columns). */
class Weighted {
    ib_weighted *h;
public:
    explicit Weighted(ib_weighted *_h) : h(_h) {}
    Weighted(Weighted const &) = delete;
    ~Weighted() { if (h) ib_w_free(h); }
    ib_weighted *handle() const { return h; }
    std::array<long, 2> shape() const { int64_t i[8]; ib_w_info(h, i); return {(long)i[0], (long)i[1]}; }
    long nnz() const { int64_t i[8]; ib_w_info(h, i); return (long)i[2]; }
    bool conservative() const { int64_t i[8]; ib_w_info(h, i); return i[3] != 0; }
    bool scaled() const { int64_t i[8]; ib_w_info(h, i); return i[4] != 0; }
    std::vector<double> get_weights(int dim) const
    {
        auto s = shape();
        std::vector<double> wM((size_t)s[0]), Mw((size_t)s[1]);
        ib_check(ib_w_get_weights(h, wM.data(), Mw.data()));
        return dim == 0 ? wM : Mw;
    }
    std::vector<int64_t> dim(int which) const { std::vector<int64_t> v((size_t)shape()[which]); ib_check(ib_w_get_dim(h, which, v.data())); return v; }
    void to_coo(std::vector<int32_t> &row, std::vector<int32_t> &col, std::vector<double> &val) const
    {
        size_t n = (size_t)nnz();
        row.resize(n); col.resize(n); val.resize(n);
        ib_check(ib_w_get_coo(h, row.data(), col.data(), val.data()));
    }
    /** apply(A_d [nvar x ncol] field-major as upstream, fill, force_conservation) -> [nvar x nrow] */
    std::vector<double> apply(std::vector<double> const &A_d, int nvar, bool force_conservation = true) const
    {
        auto s = shape();
        if ((long)A_d.size() != (long)nvar * s[1]) throw Exception(IB_ERR_DIM, "Weighted::apply: input has the wrong size");
        std::vector<double> B((size_t)nvar * s[0]);
        ib_check(ib_w_apply(h, A_d.data(), s[1], nvar, B.data(), s[0] > 0 ? s[0] : 1, IB_LAYOUT_FIELD_MAJOR, force_conservation));
        return B;
    }
    /** device-resident cell-major apply (fields never leave HBM between EvI and IvE) */
    void apply_dev(const double *dX, long ldx, int nvar, double *dY, long ldy, bool force_conservation = true, void *stream = nullptr) const
    {
        ib_check(ib_w_apply_dev(h, dX, ldx, nvar, dY, ldy, force_conservation, stream));
    }
};
/** Several matrices applied to the SAME device-resident fields in one call (the coupler's EvI and AvI on ice_ovalsI,
IceCoupler.cpp:445-446): two wide matrices are stacked and applied in one pass over X. */
inline void apply_many_dev(std::vector<Weighted const *> const &mats, const double *dX, long ldx, int nvar, std::vector<double *> const &dY,
                           std::vector<long> const &ldy, bool force_conservation = true, void *stream = nullptr)
{
    std::vector<const ib_weighted *> h;
    std::vector<int64_t> ld(ldy.begin(), ldy.end());
    for (auto m : mats) h.push_back(m->handle());
    ib_check(ib_w_apply_many_dev((int32_t)mats.size(), h.data(), dX, ldx, nvar, dY.data(), ld.data(), force_conservation, stream));
}
}   // namespace linear

/** Merge of the sheets' A- / E-grid contributions across GPUs (one rank = one process = one GPU = one sheet) over NVLink
peer memory: the device-resident replacement of "gather to the GCM root + VectorMultivec::to_dense" (modele/
GCMCoupler_ModelE.cpp:761-793, multivec.cpp:35-81).  The host supplies the one piece of communication it already has (MPI
in the reference): `allgather(send, bytes, recv)` must fill recv with every rank's `bytes` bytes in rank order.

    SheetExchange x(rank, nranks, nvar_max, allgather);
    x.add_part(gcm.nE(), *EvI_unscaled);  x.add_part(gcm.nA(), *AvI_unscaled);      // dims[0] + wM of this rank's matrices
    x.connect();
    ... every coupling step:  x.merge({dYE, dYA}, {ldE, ldA}, nvar, stream);        // unscaled sums in, merged values out
*/
class SheetExchange {
public:
    typedef std::function<void(const void *send, size_t bytes, void *recv)> AllGather;
private:
    ib_xchg *h = nullptr;
    int rank, nranks, nvar_max;
    AllGather allgather;
    struct Part { int64_t n_sparse; std::vector<int64_t> ids; std::vector<double> w; bool scaled; };
    std::vector<Part> parts;
public:
    SheetExchange(int _rank, int _nranks, int _nvar_max, AllGather const &ag) : rank(_rank), nranks(_nranks), nvar_max(_nvar_max), allgather(ag) {}
    SheetExchange(SheetExchange const &) = delete;
    ~SheetExchange() { if (h) ib_xchg_free(h); }
    /** scaled = false: the matrix is unscaled (the coupler's E1vI_unscaled) and the merge divides every row by its weight;
        scaled = true: rows are already divided by their own weight, only the shared cells are rewritten */
    void add_part(long n_sparse, linear::Weighted const &M, bool scaled = false) { parts.push_back(Part{n_sparse, M.dim(0), M.get_weights(0), scaled}); }
    void connect()
    {
        const int np = (int)parts.size();
        std::vector<int64_t> nshared((size_t)np);
        struct Plan { std::vector<int32_t> loc; std::vector<uint64_t> owners; std::vector<double> wtot; };
        std::vector<Plan> plans((size_t)np);
        for (int p = 0; p < np; ++p) {
            // gather every rank's ids and weights of this grid: counts first, then padded payloads
            int64_t mine = (int64_t)parts[p].ids.size();
            std::vector<int64_t> cnt((size_t)nranks);
            allgather(&mine, sizeof mine, cnt.data());
            int64_t mx = 0;
            for (auto c : cnt) mx = c > mx ? c : mx;
            std::vector<int64_t> sid((size_t)mx, -1), rid((size_t)mx * nranks);
            std::vector<double> sw((size_t)mx, 0.0), rw((size_t)mx * nranks);
            std::copy(parts[p].ids.begin(), parts[p].ids.end(), sid.begin());
            std::copy(parts[p].w.begin(), parts[p].w.end(), sw.begin());
            allgather(sid.data(), sizeof(int64_t) * (size_t)mx, rid.data());
            allgather(sw.data(), sizeof(double) * (size_t)mx, rw.data());
            std::vector<const int64_t *> ptr((size_t)nranks);
            for (int r = 0; r < nranks; ++r) ptr[r] = rid.data() + (size_t)r * mx;
            int64_t ns = ib_xchg_plan(nranks, ptr.data(), cnt.data(), parts[p].n_sparse, rank, nullptr, nullptr, nullptr);
            ib_check(ns);
            std::vector<int64_t> shared((size_t)ns);
            Plan &P = plans[p];
            P.loc.resize((size_t)ns); P.owners.resize((size_t)ns); P.wtot.assign((size_t)ns, 0.0);
            if (ns) ib_check(ib_xchg_plan(nranks, ptr.data(), cnt.data(), parts[p].n_sparse, rank, shared.data(), P.loc.data(), P.owners.data()));
            // summed weights of the shared cells over their owners, added in rank order
            std::map<int64_t, int64_t> where;
            for (int64_t k = 0; k < ns; ++k) where[shared[k]] = k;
            for (int r = 0; r < nranks; ++r)
                for (int64_t i = 0; i < cnt[r]; ++i) {
                    auto it = where.find(ptr[r][i]);
                    if (it != where.end()) P.wtot[(size_t)it->second] += rw[(size_t)r * mx + i];
                }
            nshared[p] = ns;
        }
        ib_check(ib_xchg_create(&h, rank, nranks, np, nshared.data(), nvar_max));
        const size_t hb = (size_t)ib_xchg_handle_bytes();
        std::vector<char> mine(hb), all(hb * nranks);
        ib_check(ib_xchg_get_handle(h, mine.data()));
        allgather(mine.data(), hb, all.data());
        ib_check(ib_xchg_open(h, all.data()));
        for (int p = 0; p < np; ++p)
            ib_check(ib_xchg_set_part(h, p, (int64_t)parts[p].ids.size(), plans[p].loc.data(), plans[p].owners.data(), plans[p].wtot.data(),
                                      parts[p].w.data(), parts[p].scaled));
    }
    void merge(std::vector<double *> const &dY, std::vector<long> const &ldy, int nvar, void *stream = nullptr)
    {
        std::vector<int64_t> ld(ldy.begin(), ldy.end());
        ib_check(ib_xchg_merge_dev(h, dY.data(), ld.data(), nvar, stream));
    }
    /** The two halves of merge() for hosts that pipeline batches: publish(s + 1) may run on another stream while gather(s)
        is still adding (each kernel keeps its own step counter); gather(s) has to follow publish(s) of the same buffers. */
    void publish(std::vector<double *> const &dY, std::vector<long> const &ldy, int nvar, void *stream = nullptr)
    {
        std::vector<int64_t> ld(ldy.begin(), ldy.end());
        ib_check(ib_xchg_publish_dev(h, dY.data(), ld.data(), nvar, stream));
    }
    void gather(std::vector<double *> const &dY, std::vector<long> const &ldy, int nvar, void *stream = nullptr)
    {
        std::vector<int64_t> ld(ldy.begin(), ldy.end());
        ib_check(ib_xchg_gather_dev(h, dY.data(), ld.data(), nvar, stream));
    }
    void status() { ib_check(ib_xchg_status(h)); }
};

class GCMRegridder_Standard;

/** Holds the set of "Ur" matrices for one ice sheet + elevmaskI; produces the regrid matrices on demand. */
class RegridMatrices_Dynamic {
    ib_rm *h;
    RegridParams _params;
public:
    RegridMatrices_Dynamic(ib_rm *_h, RegridParams const &p) : h(_h), _params(p) {}
    RegridMatrices_Dynamic(RegridMatrices_Dynamic const &) = delete;
    ~RegridMatrices_Dynamic() { if (h) ib_rm_free(h); }
    RegridParams const &params() const { return _params; }
    /** matrix(spec): dims allocated here and owned by the result (RegridMatrices_Dynamic.cpp:425-437).
        Unknown spec -> std::out_of_range, like the reference's std::map::at (:419). */
    std::unique_ptr<linear::Weighted> matrix(std::string const &spec_name) const
    {
        ib_weighted *w = nullptr;
        int st = ib_rm_matrix(h, spec_name.c_str(), &w);
        if (st == IB_ERR_ARG && std::string(ib_last_error()).rfind("map::at", 0) == 0) throw std::out_of_range(ib_last_error());
        ib_check(st);
        return std::unique_ptr<linear::Weighted>(new linear::Weighted(w));
    }
    /** matrix_d(spec, {dim0, dim1}, params): appends to the caller's dims (RegridMatrices_Dynamic.cpp:412-423) */
    std::unique_ptr<linear::Weighted> matrix_d(std::string const &spec_name, std::array<SparseSetT *, 2> dims, RegridParams const &p) const
    {
        ib_weighted *w = nullptr;
        int st = ib_rm_matrix_d(h, spec_name.c_str(), dims[0]->h, dims[1]->h, p.scale, p.correctA, p.sigma.data(), &w);
        if (st == IB_ERR_ARG && std::string(ib_last_error()).rfind("map::at", 0) == 0) throw std::out_of_range(ib_last_error());
        ib_check(st);
        return std::unique_ptr<linear::Weighted>(new linear::Weighted(w));
    }
};

struct InterpStyle { enum { Z_INTERP = IB_Z_INTERP, ELEV_CLASS_INTERP = IB_ELEV_CLASS_INTERP }; };

/** What IceRegridder::init(name, agridA, fgridA, agridI, aexgrid, interp_style) receives (IceRegridder.hpp:85-91),
    as arrays: AbbrGrid of the ice grid + ExchangeGrid + projected A areas. */
struct IceRegridder {
    std::string name;
    long nI = 0;                             // sparse extent of the ice grid
    std::vector<int64_t> dimI;               // realised ice cells (dense -> sparse)
    std::vector<double> native_areaI;        // per realised cell
    std::vector<double> centroid_xy;         // [n][2], may be empty (then no smoothing)
    std::vector<int32_t> ex_ij;              // ExchangeGrid::indices, interleaved (iA, iI)
    std::vector<double> ex_area;             // ExchangeGrid::overlaps
    std::vector<double> gridA_proj_area;     // per realised A cell
    int interp_style = InterpStyle::Z_INTERP;
    long nX() const { return (long)ex_area.size(); }
};

class GCMRegridder_Standard {
    ib_gcm *h = nullptr;
    std::map<std::string, int> sheets_index;
    size_t n_real = 0;
public:
    GCMRegridder_Standard() {}
    GCMRegridder_Standard(GCMRegridder_Standard const &) = delete;
    ~GCMRegridder_Standard() { if (h) ib_gcm_free(h); }
    /** init(agridA, hcdefs, indexingHC, correctA) with indexingHC = {A,HC}, HC slowest (icebin_cython.cpp:56-73) */
    void init(long nA, std::vector<int64_t> const &dimA, std::vector<double> const &native_area, std::vector<double> const &hcdefs, bool correctA)
    {
        if (h) { ib_gcm_free(h); h = nullptr; }
        n_real = dimA.size();
        ib_check(ib_gcm_create(&h, nA, (int32_t)dimA.size(), dimA.data(), native_area.data(), (int32_t)hcdefs.size(), hcdefs.data(), correctA));
    }
    int add_sheet(IceRegridder const &s)
    {
        if (s.gridA_proj_area.size() != n_real) throw Exception(IB_ERR_DIM, "add_sheet: gridA_proj_area must have one entry per realised A cell");
        int ix = ib_gcm_add_sheet(h, s.name.c_str(), s.nI, (int32_t)s.dimI.size(), s.dimI.data(), s.native_areaI.data(),
                                  s.centroid_xy.empty() ? nullptr : s.centroid_xy.data(), s.nX(), s.ex_ij.data(), s.ex_area.data(),
                                  s.gridA_proj_area.data(), s.interp_style);
        ib_check(ix);
        sheets_index[s.name] = ix;
        return ix;
    }
    /** filter_cellsA(keepA) (GCMRegridder.hpp:380-389, GCMRegridder.cpp:154-164): keep only the GCM cells of this rank's /
        GPU's domain; the exchange grids of all sheets are compacted on the device */
    void filter_cellsA(std::function<bool(long)> const &keepA)
    {
        std::vector<uint8_t> keep((size_t)nA());
        for (size_t i = 0; i < keep.size(); ++i) keep[i] = keepA((long)i) ? 1 : 0;
        ib_check(ib_gcm_filter_cellsA(h, keep.data()));
        n_real = (size_t)ib_gcm_nA_realized(h);
    }
    long nX(int sheet_index) const { return (long)ib_gcm_nX(h, sheet_index); }
    unsigned long nA() const { return (unsigned long)ib_gcm_nA(h); }
    unsigned long nE() const { return (unsigned long)ib_gcm_nE(h); }
    unsigned int nhc() const { return (unsigned int)ib_gcm_nhc(h); }
    size_t nI(int sheet_index) const { return (size_t)ib_gcm_nI(h, sheet_index); }
    int sheet_index(std::string const &name) const { return sheets_index.at(name); }
    /** wA(accum, sheet, native): (sparse A index, area) pairs (GCMRegridder.hpp:304-314) */
    void wA(std::vector<int64_t> &index, std::vector<double> &area, std::string const &sheet, bool native) const
    {
        index.resize(n_real); area.resize(n_real);
        ib_check(ib_gcm_wA(h, sheets_index.at(sheet), native, index.data(), area.data()));
    }
    /** regrid_matrices(sheet_index, elevmaskI, params): elevmaskI is copied (RegridMatrices_Dynamic.cpp:334-402) */
    std::unique_ptr<RegridMatrices_Dynamic> regrid_matrices(int sheet_index, std::vector<double> const &elevmaskI, RegridParams const &params = RegridParams()) const
    {
        if (elevmaskI.size() != nI(sheet_index)) throw Exception(IB_ERR_DIM, "regrid_matrices: elevmaskI must have nI elements");
        ib_rm *r = nullptr;
        ib_check(ib_regrid_matrices(h, sheet_index, elevmaskI.data(), params.scale, params.correctA, params.sigma.data(), &r));
        return std::unique_ptr<RegridMatrices_Dynamic>(new RegridMatrices_Dynamic(r, params));
    }
};

/** What combine_chunks() of modele/combine_global_ec.cpp:95-262 accumulates into its Weighted_Compressed: the chunks'
    entries relabelled with sparse indices (scaled by 1 / sum of wM over the chunks if `scale`), and wM / Mw as sparse lists */
struct CombinedChunks {
    std::vector<int64_t> M_row, M_col, wM_index, Mw_index;
    std::vector<double> M_val, wM_val, Mw_val;
};
inline CombinedChunks combine_chunks(std::vector<linear::Weighted const *> const &chunks, bool scale)
{
    CombinedChunks c;
    size_t nnz = 0, nr = 0, nc = 0;
    std::vector<const ib_weighted *> w;
    for (auto *ch : chunks) { nnz += (size_t)ch->nnz(); nr += (size_t)ch->shape()[0]; nc += (size_t)ch->shape()[1]; w.push_back(ch->handle()); }
    c.M_row.resize(nnz); c.M_col.resize(nnz); c.M_val.resize(nnz);
    c.wM_index.resize(nr); c.wM_val.resize(nr); c.Mw_index.resize(nc); c.Mw_val.resize(nc);
    ib_check(ib_w_combine_chunks((int32_t)w.size(), w.data(), scale, c.M_row.data(), c.M_col.data(), c.M_val.data(), c.wM_index.data(),
                                 c.wM_val.data(), c.Mw_index.data(), c.Mw_val.data()));
    return c;
}

/** make_exchange_grid (gridgen/GridGen_Exchange.cpp:208-316) for GCM cells given as polygons projected into the plane of a
    rectilinear ice grid: fills ExchangeGrid::indices (interleaved iA, iI) and ::overlaps (AbbrGrid.hpp:40-78) */
inline void make_exchange_grid_rectilinear(std::vector<int32_t> const &pptr, std::vector<double> const &pxy, std::vector<int64_t> const &iA,
                                           std::vector<double> const &xe, std::vector<double> const &ye, bool iI_major_x,
                                           std::vector<int32_t> &indices, std::vector<double> &overlaps)
{
    const int32_t npoly = (int32_t)iA.size();
    int64_t n = ib_exgrid_rectilinear(npoly, pptr.data(), pxy.data(), iA.data(), (int32_t)xe.size() - 1, xe.data(), (int32_t)ye.size() - 1,
                                      ye.data(), iI_major_x, nullptr, nullptr);
    ib_check(n);
    indices.resize((size_t)(2 * n)); overlaps.resize((size_t)n);
    if (n > 0)
        ib_check(ib_exgrid_rectilinear(npoly, pptr.data(), pxy.data(), iA.data(), (int32_t)xe.size() - 1, xe.data(), (int32_t)ye.size() - 1,
                                       ye.data(), iI_major_x, indices.data(), overlaps.data()));
}

namespace modele {

/** im, jm, offi (cells from the date line to the western edge of cell 1), dlat in MINUTES (GridSpec.hpp:143-170) */
struct HntrSpec {
    int im, jm;
    double offi, dlat;
    HntrSpec() : im(-1), jm(-1), offi(0.), dlat(0.) {}
    HntrSpec(int _im, int _jm, double _offi, double _dlat) : im(_im), jm(_jm), offi(_offi), dlat(_dlat) {}
    int size() const { return im * jm; }
    int ndata() const { return size(); }
    bool is_set() const { return im >= 0; }
};

class Hntr {
    ib_hntr *h = nullptr;
public:
    HntrSpec const Agrid, Bgrid;
    /** Hntr(yp17, Bspec, Aspec, DATMIS): yp17 is ignored, as upstream (hntr.cpp:63) */
    Hntr(double /*yp17*/, HntrSpec const &_B, HntrSpec const &_A, double DATMIS = 0.0) : Agrid(_A), Bgrid(_B)
    {
        ib_check(ib_hntr_create(&h, _B.im, _B.jm, _B.offi, _B.dlat, _A.im, _A.jm, _A.offi, _A.dlat, DATMIS));
    }
    Hntr(Hntr const &) = delete;
    ~Hntr() { if (h) ib_hntr_free(h); }
    /** regrid(WTA, A, B, mean_polar, wtm, wtb) (hntr.hpp:385-435); WTA == nullptr is a constant weight of 1 */
    void regrid(const double *WTA, const double *A, double *B, bool mean_polar = false, double wtm = 1.0, double wtb = 0.0) const
    {
        ib_check(ib_hntr_regrid(h, WTA, wtm, wtb, A, IB_SRC_F64, B, mean_polar));
    }
    void regrid(const double *WTA, const int16_t *A, double *B, bool mean_polar = false, double wtm = 1.0, double wtb = 0.0) const
    {
        ib_check(ib_hntr_regrid(h, WTA, wtm, wtb, A, IB_SRC_I16, B, mean_polar));
    }
    /** overlap(accum, R) / scaled_regrid_matrix(accum) (hntr.hpp:249-339): 0-based (IJB-1, IJA-1, value) triplets */
    void overlap(std::vector<int64_t> &rows, std::vector<int64_t> &cols, std::vector<double> &vals, double R) const { matrix(rows, cols, vals, R); }
    void scaled_regrid_matrix(std::vector<int64_t> &rows, std::vector<int64_t> &cols, std::vector<double> &vals) const { matrix(rows, cols, vals, -1.0); }
private:
    void matrix(std::vector<int64_t> &rows, std::vector<int64_t> &cols, std::vector<double> &vals, double R) const
    {
        int64_t n = ib_hntr_matrix(h, R, nullptr, nullptr, nullptr);
        ib_check(n);
        rows.resize((size_t)n); cols.resize((size_t)n); vals.resize((size_t)n);
        ib_check(ib_hntr_matrix(h, R, rows.data(), cols.data(), vals.data()));
    }
};

}   // namespace modele
}   // namespace icebin
