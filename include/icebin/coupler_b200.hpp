// coupler_b200.hpp -- host-side C++ mirror of the REGRIDDING side of the reference's coupling step, over the C ABI.
//
//   icebin::VectorMultivec            slib/icebin/multivec.hpp (index / weights / vals triples crossing to the GCM)
//   icebin::VarTransT                 what VarTransformer::apply_scalars(scalars, 'T') returns: sparse M + row vector b
//   icebin::IceCouplerRegrid          IceCoupler::construct_ice_ivalsI (IceCoupler.cpp:203-252) and the part of
//                                     IceCoupler::couple that rebuilds the matrices and regrids the ice model's output
//                                     (IceCoupler.cpp:340-497), with the state it keeps between steps (dimE0, IvE0)
//   icebin::e1ve0::compute_E1vE0c     e1ve0.cpp:55-106 (call site GCMCoupler.cpp:488-515)
// The ice model itself, contracts, NetCDF logging and MPI are the caller's (out of scope).
// Fields are [nvar x ncell] (variable slowest) on the host, exactly like the reference's Blitz arrays
// (IceCoupler.cpp:231-233); on the device they are cell-major and stay in HBM between the variable transformation and
// the regrid applies.  Header-only; link with -licebin_b200.
#pragma once
#include "regrid_b200.hpp"

namespace icebin {

/** multivec.hpp: sparse multi-variable vector, appended to by every ice sheet */
struct VectorMultivec {
    int nvar;
    std::vector<long> index;
    std::vector<double> weights;
    std::vector<double> vals;      // [index.size()][nvar]
    explicit VectorMultivec(int _nvar) : nvar(_nvar) {}
    size_t size() const { return index.size(); }
    double val(int ivar, size_t i) const { return vals[i * nvar + ivar]; }
    void add(long ix, std::vector<double> const &v, double weight)
    {
        index.push_back(ix);
        weights.push_back(weight);
        vals.insert(vals.end(), v.begin(), v.begin() + nvar);
    }
    /** to_dense_scale + to_dense for every variable (multivec.cpp:35-81): dense[nvar][nE] */
    void to_dense(long nE, double fill, std::vector<double> &scaleE, std::vector<double> &dense) const
    {
        std::vector<int64_t> ix(index.begin(), index.end());
        scaleE.assign((size_t)nE, 0.);
        dense.assign((size_t)nE * nvar, 0.);
        ib_check(ib_multivec_to_dense((int64_t)ix.size(), ix.data(), weights.data(), vals.data(), nvar, nE, fill, scaleE.data(), dense.data()));
    }
};

/** Transposed variable transformation M(input, output) + row vector b, column-major sparse as Eigen holds it */
struct VarTransT {
    int nin = 0, nout = 0;
    std::vector<int32_t> colptr, row;
    std::vector<double> val, b;
    /** identity transformation of n variables */
    static VarTransT identity(int n)
    {
        VarTransT t;
        t.nin = t.nout = n;
        for (int k = 0; k <= n; ++k) t.colptr.push_back(k);
        for (int k = 0; k < n; ++k) { t.row.push_back(k); t.val.push_back(1.0); t.b.push_back(0.0); }
        return t;
    }
};

namespace detail {
struct DevBuf {
    void *p = nullptr;
    explicit DevBuf(size_t ndouble) { ib_check(ib_dev_alloc(&p, (int64_t)(ndouble * sizeof(double)))); }
    DevBuf(DevBuf const &) = delete;
    ~DevBuf() { ib_dev_free(p); }
    double *d() const { return (double *)p; }
};
/** host [nvar x ncell] -> device [ncell x nvar] */
inline void upload_cell_major(DevBuf &dst, std::vector<double> const &src, int nvar, long ncell)
{
    std::vector<double> t((size_t)nvar * ncell);
    for (int v = 0; v < nvar; ++v)
        for (long i = 0; i < ncell; ++i) t[(size_t)i * nvar + v] = src[(size_t)v * ncell + i];
    ib_check(ib_dev_upload(dst.p, t.data(), (int64_t)(t.size() * sizeof(double))));
}
inline void vartrans(DevBuf const &O, long ncell, VarTransT const &t, DevBuf &Z)
{
    ib_check(ib_vartrans_dev(O.d(), t.nin, ncell, t.nin, t.nout, t.colptr.data(), t.row.data(), t.val.data(), t.b.data(), Z.d(), t.nout, nullptr));
}
}   // namespace detail

class IceCouplerRegrid {
    GCMRegridder_Standard const *gcmr;
    int sheet_index;
    long _nI, _nX;
    std::array<double, 3> sigma;
    // state carried from one coupling step to the next (IceCoupler.hpp: dimE0, IvE0)
    std::unique_ptr<SparseSetT> dimE0;
    std::unique_ptr<linear::Weighted> IvE0;
public:
    struct CoupleOut {
        // matrix_d() borrows its dims (RegridMatrices_Dynamic.cpp:412-423); the reference sparsifies XuE at once
        // (GCMCoupler.cpp:500-505) -- here XuE travels with dims of its own, declared first so that they outlive it
        std::unique_ptr<SparseSetT> dimX, dimE_XuE;
        std::unique_ptr<linear::Weighted> XuE;      // "XvE", scale=f, correctA=t: input of compute_E1vE0c
        SparseSetT const *dimE = nullptr;           // elevation classes in use after this step (owned by the coupler)
    };
    IceCouplerRegrid(GCMRegridder_Standard const *_gcmr, int _sheet_index, long nX, std::array<double, 3> const &_sigma = {0., 0., 0.})
        : gcmr(_gcmr), sheet_index(_sheet_index), _nI((long)_gcmr->nI(_sheet_index)), _nX(nX), sigma(_sigma) {}
    long nI() const { return _nI; }
    SparseSetT const *dimE() const { return dimE0.get(); }

    /** ice_ivalsI = IvE0 * (gcm_ovalsE0 * ivg.M + ivg.b)   IceCoupler.cpp:203-252.
        gcm_ovalsE0: [ivg.nin x dimE0.dense_extent()] densified on dimE0 (:296-307).  Returns [ivg.nout x nI]. */
    std::vector<double> construct_ice_ivalsI(std::vector<double> const &gcm_ovalsE0, VarTransT const &ivg) const
    {
        if (!IvE0) throw Exception(IB_ERR_ARG, "construct_ice_ivalsI: no IvE0 yet (couple_regrid() has not run)");
        const long nE0 = IvE0->shape()[1];
        if ((long)gcm_ovalsE0.size() != (long)ivg.nin * nE0) throw Exception(IB_ERR_DIM, "construct_ice_ivalsI: gcm_ovalsE0 has the wrong size");
        detail::DevBuf O((size_t)ivg.nin * nE0), Z((size_t)ivg.nout * nE0), Y((size_t)ivg.nout * _nI);
        detail::upload_cell_major(O, gcm_ovalsE0, ivg.nin, nE0);
        detail::vartrans(O, nE0, ivg, Z);
        // the explicit Eigen product of the reference: no conservation correction (IceCoupler.cpp:237-238)
        IvE0->apply_dev(Z.d(), ivg.nout, ivg.nout, Y.d(), ivg.nout, false);
        std::vector<double> t((size_t)ivg.nout * _nI), out(t.size());
        ib_check(ib_dev_download(t.data(), Y.p, (int64_t)(t.size() * sizeof(double))));
        for (long i = 0; i < _nI; ++i)
            for (int v = 0; v < ivg.nout; ++v) out[(size_t)v * _nI + i] = t[(size_t)i * ivg.nout + v];
        return out;
    }

    /** The regridding part of IceCoupler::couple (IceCoupler.cpp:340-497): rebuild the matrices for the ice model's
        new elevation mask, regrid ice_ovalsI [nvar x nI] to the A and E grids through the variable transformations
        gvi[A], gvi[E], append (sparse index, values, weight wM) to gcm_ivalss_s[A], [E], and keep IvE / dimE for the next
        step.  emI_ice: elevmask_ice [nI] (NaN = no ice). */
    CoupleOut couple_regrid(std::vector<double> const &ice_ovalsI, int nvar, std::vector<double> const &emI_ice,
                            std::array<VarTransT const *, 2> gvi, std::array<VectorMultivec *, 2> gcm_ivalss_s)
    {
        if ((long)ice_ovalsI.size() != (long)nvar * _nI) throw Exception(IB_ERR_DIM, "couple_regrid: ice_ovalsI has the wrong size");
        auto rm = gcmr->regrid_matrices(sheet_index, emI_ice);
        SparseSetT dimI = SparseSetT::identity(_nI);
        std::unique_ptr<SparseSetT> dimE1(new SparseSetT);
        SparseSetT dimA1;
        // _nc: no correction for projection area changes (IceCoupler.cpp:368-372); scale=f throughout
        auto E1vI_unscaled_nc = rm->matrix_d("EvI", {dimE1.get(), &dimI}, RegridParams(false, false));
        auto A1vI_unscaled = rm->matrix_d("AvI", {&dimA1, &dimI}, RegridParams(false, true));
        linear::Weighted *AE1vIs[2] = {A1vI_unscaled.get(), E1vI_unscaled_nc.get()};

        detail::DevBuf O((size_t)nvar * _nI);
        detail::upload_cell_major(O, ice_ovalsI, nvar, _nI);
        for (int iAE = 0; iAE < 2; ++iAE) {
            VarTransT const &t = *gvi[iAE];
            if (t.nin != nvar || t.nout != gcm_ivalss_s[iAE]->nvar) throw Exception(IB_ERR_DIM, "couple_regrid: variable transformation does not match the contracts");
            linear::Weighted &W = *AE1vIs[iAE];
            const long nrow = W.shape()[0];
            detail::DevBuf Z((size_t)t.nout * _nI), Y((size_t)t.nout * (nrow > 0 ? nrow : 1));
            detail::vartrans(O, _nI, t, Z);
            W.apply_dev(Z.d(), t.nout, t.nout, Y.d(), t.nout, false);
            std::vector<double> y((size_t)t.nout * nrow);
            ib_check(ib_dev_download(y.data(), Y.p, (int64_t)(y.size() * sizeof(double))));
            // sparsify while appending to the global VectorMultivec (IceCoupler.cpp:455-466)
            auto ids = W.dim(0);
            auto wM = W.get_weights(0);
            std::vector<double> vals((size_t)t.nout);
            for (long jj = 0; jj < nrow; ++jj) {
                for (int nn = 0; nn < t.nout; ++nn) vals[nn] = y[(size_t)jj * t.nout + nn];
                gcm_ivalss_s[iAE]->add((long)ids[jj], vals, wM[jj]);
            }
        }
        // IvE for interpreting the GCM's E-grid fields at the start of the next step (:470-472) and XuE (:474-477)
        auto IvE1 = rm->matrix_d("IvE", {&dimI, dimE1.get()}, RegridParams(true, true, sigma));
        CoupleOut ret;
        ret.dimX.reset(new SparseSetT(SparseSetT::identity(_nX)));
        ret.dimE_XuE.reset(new SparseSetT((long)gcmr->nE(), dimE1->to_sparse()));       // same ids, same order as dimE1
        ret.XuE = rm->matrix_d("XvE", {ret.dimX.get(), ret.dimE_XuE.get()}, RegridParams(false, true));
        dimE0 = std::move(dimE1);
        IvE0 = std::move(IvE1);
        ret.dimE = dimE0.get();
        return ret;
    }
};

namespace e1ve0 {
struct Tuple { long iE1, iE0; double value; };
/** compute_E1vE0c(XuE1s, XuE0s, nE, areaX)  e1ve0.cpp:55-106 (areaX is unused upstream as well) */
inline std::vector<Tuple> compute_E1vE0c(std::vector<std::unique_ptr<linear::Weighted>> const &XuE1s,
                                         std::vector<std::unique_ptr<linear::Weighted>> const &XuE0s, unsigned long nE)
{
    if (XuE1s.size() != XuE0s.size() || XuE1s.empty()) throw Exception(IB_ERR_ARG, "compute_E1vE0c: one XuE pair per ice sheet");
    std::vector<const ib_weighted *> a, b;
    for (size_t i = 0; i < XuE1s.size(); ++i) { a.push_back(XuE1s[i]->handle()); b.push_back(XuE0s[i]->handle()); }
    int64_t n = ib_e1ve0c((int32_t)a.size(), a.data(), b.data(), (int64_t)nE, nullptr, nullptr, nullptr);
    ib_check(n);
    std::vector<int64_t> r((size_t)n), c((size_t)n);
    std::vector<double> v((size_t)n);
    ib_check(ib_e1ve0c((int32_t)a.size(), a.data(), b.data(), (int64_t)nE, r.data(), c.data(), v.data()));
    std::vector<Tuple> out((size_t)n);
    for (int64_t i = 0; i < n; ++i) out[(size_t)i] = Tuple{(long)r[i], (long)c[i], v[i]};
    return out;
}
}   // namespace e1ve0

}   // namespace icebin
