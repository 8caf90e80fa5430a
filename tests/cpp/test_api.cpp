// Exercises the C++ host API (include/icebin/regrid_b200.hpp) the way the reference's C++ callers use
// GCMRegridder_Standard / RegridMatrices_Dynamic / Hntr (IceCoupler.cpp:361-468, modele/etopo1_ice.cpp:17-29).
// Built by tests/test_cpp_api.py with g++ (no nvcc needed); needs a GPU to RUN ("--compile-only" just links).
#include <cstdio>
#include <cstring>

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

    std::printf("C++ API ok\n");
    return 0;
}
