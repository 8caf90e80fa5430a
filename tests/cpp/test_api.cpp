// Exercises the C++ host API (include/icebin/regrid_b200.hpp) the way the reference's C++ callers use
// GCMRegridder_Standard / RegridMatrices_Dynamic / Hntr (IceCoupler.cpp:361-468, modele/etopo1_ice.cpp:17-29).
// Built by tests/test_cpp_api.py with g++ (no nvcc needed); needs a GPU to RUN ("--compile-only" just links).
#include <cstdio>
#include <cstring>

#include "icebin/coupler_b200.hpp"
#include "icebin/regrid_b200.hpp"

using namespace icebin;

#define CHECK(c)                                                            \
    do {                                                                    \
        if (!(c)) { std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); return 1; } \
    } while (0)

int main(int argc, char **argv)
{
    if (argc > 1 && !std::strcmp(argv[1], "--compile-only")) { std::printf("linked ok, %d device(s)\n", ib_device_count()); return 0; }

    // A grid: 4 x 3 cells; ice grid 8 x 6 cells of size 1, each A cell covers 2 x 2 ice cells (exact nesting)
    const long nA = 12, nI = 48;
    std::vector<int64_t> dimA(nA);
    std::vector<double> areaA(nA);
    for (long a = 0; a < nA; ++a) { dimA[a] = a; areaA[a] = 4.0 * (1.0 + 0.01 * a); }
    std::vector<double> hcdefs = {-50, 50, 150, 250};
    GCMRegridder_Standard gcm;
    gcm.init(nA, dimA, areaA, hcdefs, true);

    IceRegridder sheet;
    sheet.name = "greenland";
    sheet.nI = nI;
    std::vector<double> em(nI);
    for (long ix = 0; ix < 8; ++ix)
        for (long iy = 0; iy < 6; ++iy) {
            long iI = ix * 6 + iy;
            sheet.dimI.push_back(iI);
            sheet.native_areaI.push_back(1.0);
            sheet.centroid_xy.push_back(ix + .5);
            sheet.centroid_xy.push_back(iy + .5);
            long iA = (iy / 2) * 4 + (ix / 2);
            sheet.ex_ij.push_back((int32_t)iA);
            sheet.ex_ij.push_back((int32_t)iI);
            sheet.ex_area.push_back(1.0);
            em[iI] = (ix == 0) ? std::nan("") : 10.0 * ix + 7.0 * iy;    // first column masked out
        }
    sheet.gridA_proj_area.assign(nA, 4.0);
    CHECK(gcm.add_sheet(sheet) == 0);
    CHECK(gcm.nA() == 12 && gcm.nE() == 48 && gcm.nhc() == 4 && gcm.nI(0) == 48);

    auto rm = gcm.regrid_matrices(0, em, RegridParams(true, true));
    auto EvI = rm->matrix("EvI");
    auto IvE = rm->matrix("IvE");
    auto AvI = rm->matrix("AvI");
    CHECK(AvI->shape()[1] == 42 && AvI->nnz() == 42);          // 48 - 6 masked ice cells
    CHECK(EvI->conservative() && EvI->scaled());
    CHECK(EvI->dim(1) == IvE->dim(0) && EvI->dim(0) == IvE->dim(1));

    // EvI and AvI on the same device-resident fields in one call (IceCoupler.cpp:445-446), then the single-rank form of the
    // multi-GPU merge: with one rank no cell is shared, so unscaled sums / wM must equal the scaled matrices' results
    {
        auto rmu = gcm.regrid_matrices(0, em, RegridParams(false, true));
        auto EvIu = rmu->matrix("EvI");
        auto AvIu = rmu->matrix("AvI");
        const int nv = 64;
        const long nc = EvI->shape()[1], nE2 = EvI->shape()[0], nA2 = AvI->shape()[0];
        std::vector<double> X((size_t)nc * nv);
        for (size_t k = 0; k < X.size(); ++k) X[k] = 0.25 + 0.001 * (double)(k % 977);
        detail::DevBuf dX(X.size()), dE((size_t)nE2 * nv), dA((size_t)nA2 * nv), dE2((size_t)nE2 * nv), dA2((size_t)nA2 * nv);
        ib_check(ib_dev_upload(dX.p, X.data(), (int64_t)(X.size() * sizeof(double))));
        linear::apply_many_dev({EvI.get(), AvI.get()}, dX.d(), nv, nv, {dE.d(), dA.d()}, {nv, nv});
        linear::apply_many_dev({EvIu.get(), AvIu.get()}, dX.d(), nv, nv, {dE2.d(), dA2.d()}, {nv, nv});
        SheetExchange xch(0, 1, nv, [](const void *send, size_t bytes, void *recv) { std::memcpy(recv, send, bytes); });
        xch.add_part(gcm.nE(), *EvIu);
        xch.add_part(gcm.nA(), *AvIu);
        xch.connect();
        xch.merge({dE2.d(), dA2.d()}, {nv, nv}, nv);
        xch.status();
        std::vector<double> a((size_t)nE2 * nv), b((size_t)nE2 * nv);
        ib_check(ib_dev_download(a.data(), dE.p, (int64_t)(a.size() * sizeof(double))));
        ib_check(ib_dev_download(b.data(), dE2.p, (int64_t)(b.size() * sizeof(double))));
        for (size_t k = 0; k < a.size(); ++k) CHECK(std::fabs(a[k] - b[k]) <= 1e-12 * std::fabs(a[k]));
    }

    // conservation: sum wM * (EvI f) == sum Mw * f
    long nId = EvI->shape()[1], nEd = EvI->shape()[0];
    std::vector<double> f((size_t)nId);
    for (long i = 0; i < nId; ++i) f[i] = 1.0 + 0.1 * i;
    auto y = EvI->apply(f, 1);
    auto wM = EvI->get_weights(0), Mw = EvI->get_weights(1);
    double lhs = 0, rhs = 0;
    for (long e = 0; e < nEd; ++e) lhs += wM[e] * y[e];
    for (long i = 0; i < nId; ++i) rhs += Mw[i] * f[i];
    CHECK(std::abs(lhs / rhs - 1) < 1e-12);

    // matrix_d with caller dims, IceCoupler style (scale=false), shares dimE between two matrices
    SparseSetT dimE, dimI = SparseSetT::identity(nI);
    auto EvI_u = rm->matrix_d("EvI", {&dimE, &dimI}, RegridParams(false, false));
    long nE1 = dimE.dense_extent();
    auto IvE_s = rm->matrix_d("IvE", {&dimI, &dimE}, RegridParams(true, true));
    CHECK(dimE.dense_extent() == nE1 && EvI_u->shape()[1] == nI && !EvI_u->scaled());

    bool threw = false;
    try { rm->matrix("IvI"); } catch (std::out_of_range const &) { threw = true; }
    CHECK(threw);

    // Hntr: 16x8 -> 8x4 of a constant is the constant; integral conserved
    modele::HntrSpec g4(8, 4, 0.0, 45.0 * 60), g8(16, 8, 0.0, 22.5 * 60);
    modele::Hntr hntr(17.17, g4, g8);
    std::vector<double> A(128, 3.25), B(32);
    hntr.regrid(nullptr, A.data(), B.data());
    for (double b : B) CHECK(std::abs(b - 3.25) < 1e-14);
    std::vector<int64_t> r, c;
    std::vector<double> v;
    hntr.overlap(r, c, v, 1.0);
    double tot = 0;
    for (double x : v) tot += x;
    CHECK(std::abs(tot / (4 * M_PI) - 1) < 1e-12);
    threw = false;
    try { modele::Hntr bad(17.17, modele::HntrSpec(144, 90, 0., 2.0), g8); } catch (Exception const &) { threw = true; }
    CHECK(threw);

    // coupling step, regridding side (IceCoupler.cpp:340-497, :203-252, e1ve0.cpp:55-106)
    {
        IceCouplerRegrid cpl(&gcm, 0, sheet.nX());
        const int nv = 2;
        std::vector<double> ice_ovalsI((size_t)nv * nI);
        for (long i = 0; i < nI; ++i) { ice_ovalsI[i] = 1.0; ice_ovalsI[nI + i] = 0.5 * i; }
        VarTransT id2 = VarTransT::identity(nv);
        id2.b[0] = 2.0;                                               // first variable: x -> x + 2
        VectorMultivec outA(nv), outE(nv);
        auto step0 = cpl.couple_regrid(ice_ovalsI, nv, em, {&id2, &id2}, {&outA, &outE});
        CHECK(outE.size() == (size_t)cpl.dimE()->dense_extent() && outA.size() > 0);
        // E side is built without area correction: regridding the constant 1 (+2) gives weight * 3
        for (size_t i = 0; i < outE.size(); ++i) CHECK(std::abs(outE.val(0, i) / outE.weights[i] - 3.0) < 1e-12);
        std::vector<double> scaleE, dense;
        outE.to_dense((long)gcm.nE(), -1.0, scaleE, dense);
        long touched = 0;
        for (unsigned long e = 0; e < gcm.nE(); ++e) { if (dense[e] != -1.0) { ++touched; CHECK(std::abs(dense[e] - 3.0) < 1e-12); } }
        CHECK(touched == (long)outE.size());
        // gcm -> ice with the IvE of that step (scale=t, correctA=t): a constant E field comes back as the constant times the
        // native/projected area ratio of the GCM cell the ice cell lies in (M = diag(1/wM) IvEp diag(sEpvE))
        const long nE0 = cpl.dimE()->dense_extent();
        std::vector<double> gcm_ovalsE0((size_t)nv * nE0, 4.0);
        auto ice_ivalsI = cpl.construct_ice_ivalsI(gcm_ovalsE0, id2);
        for (long i = 0; i < nI; ++i) {
            if (std::isnan(em[i])) continue;
            const long ix = i / 6, iy = i % 6, iA = (iy / 2) * 4 + (ix / 2);
            const double ratio = areaA[iA] / 4.0;
            CHECK(std::abs(ice_ivalsI[i] / (6.0 * ratio) - 1) < 1e-12 && std::abs(ice_ivalsI[nI + i] / (4.0 * ratio) - 1) < 1e-12);
        }
        // second step with a thinner ice sheet: E1vE0c relates the two sets of elevation classes
        std::vector<double> em1(em);
        for (double &e : em1) if (!std::isnan(e)) e = 0.6 * e;
        VectorMultivec outA1(nv), outE1(nv);
        auto step1 = cpl.couple_regrid(ice_ovalsI, nv, em1, {&id2, &id2}, {&outA1, &outE1});
        std::vector<std::unique_ptr<linear::Weighted>> X1, X0;
        X1.push_back(std::move(step1.XuE));
        X0.push_back(std::move(step0.XuE));
        auto c10 = e1ve0::compute_E1vE0c(X1, X0, gcm.nE());
        CHECK(!c10.empty());
        bool nonzero = false;
        for (size_t i = 1; i < c10.size(); ++i) CHECK(c10[i - 1].iE1 < c10[i].iE1 || (c10[i - 1].iE1 == c10[i].iE1 && c10[i - 1].iE0 < c10[i].iE0));
        for (auto const &t : c10) nonzero = nonzero || t.value != 0.0;
        CHECK(nonzero);
        auto c11 = e1ve0::compute_E1vE0c(X1, X1, gcm.nE());
        for (auto const &t : c11) CHECK(t.value == 0.0);
    }

    std::printf("C++ API ok\n");
    return 0;
}
