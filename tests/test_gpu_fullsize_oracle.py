"""GPU vs the CPU oracle at BASELINE.json's FULL sizes (configs[3] and configs[4]); the C++ oracle needs seconds here.
  * Hntr 21600x10800 -> 144x90 and 360x180, int16 source / constant weight and fp64 source / fp64 weights, mean_polar off
    and on, every output cell against orc.Hntr to 1e-12          (tests/test_hntr.cpp:174-184, 308, 372-399);
  * Greenland 1 km (1501x2801) x 40 ECs: the six matrices against the oracle -- dense<->sparse maps and sparsity pattern
    bit for bit, values and weights to 1e-12                     (tests/test_conserv/test_conserv.py:89-205);
  * configs[4] proper: Greenland 1 km + Antarctica 1 km on ONE GCMRegridder; every sheet's unscaled EvI / AvI applied on
    the GPU, contributions merged on the A and E grids, against the oracle's matrices and orc.multivec_to_dense
    (slib/icebin/multivec.cpp:35-81)."""
import numpy as np
import pytest
import torch

from conftest import make_gpu, make_oracle
from icebin_b200 import api, synth
from oracle import orc

pytestmark = pytest.mark.gpu

SPEC = {"g2hx2": (144, 90, 0.0, 120.0), "g1x1": (360, 180, 0.0, 60.0), "g1mx1m": (21600, 10800, 0.0, 1.0)}


def rel_close(a, b, rtol, what):
    scale = np.maximum(np.abs(b), np.abs(b).max() * 1e-3)      # cells near zero are held to the field's scale
    bad = ~(np.abs(a - b) <= rtol * scale)
    assert not bad.any(), "%s: %d of %d cells differ, worst %.3e relative" % (what, bad.sum(), bad.size, (np.abs(a - b) / scale).max())


@pytest.mark.parametrize("bname", ["g2hx2", "g1x1"])
def test_hntr_etopo1_against_oracle(bname):
    A, B = api.GRIDS["g1mx1m"], api.GRIDS[bname]
    assert (A.im, A.jm, B.im, B.jm) == SPEC["g1mx1m"][:2] + SPEC[bname][:2]
    rng = np.random.default_rng(4)
    hg = api.Hntr(17.17, B, A)
    ho = orc.Hntr(SPEC[bname], SPEC["g1mx1m"])
    # etopo1_ice.cpp:17-29: int16 topography, constant weight
    src16 = rng.integers(-500, 4000, A.size, dtype=np.int16)
    for mean_polar in (False, True):
        rel_close(hg.regrid(None, src16, mean_polar=mean_polar), ho.regrid(None, src16, mean_polar=mean_polar), 1e-12,
                  "%s int16 const-weight mean_polar=%s" % (bname, mean_polar))
    # fp64 source with a 30 % mask as weights (the generic path; destination cells without weight keep DATMIS)
    src64 = rng.uniform(-1.0, 1.0, A.size)
    w64 = (rng.random(A.size) < 0.3).astype(np.float64)
    for mean_polar in (False, True):
        rel_close(hg.regrid(w64, src64, mean_polar=mean_polar), ho.regrid(w64, src64, mean_polar=mean_polar), 1e-12,
                  "%s fp64 fp64-weights mean_polar=%s" % (bname, mean_polar))


def sorted_coo(Wg):
    r, c, v = Wg.to_coo_dense()
    o = np.lexsort((r, c))            # the oracle's (column-major) order
    return r[o], c[o], v[o]


@pytest.fixture(scope="module")
def cfg5():
    return synth.config(5)


def one_sheet(cfg5, k):
    return dict(gridA=cfg5["gridA"], sheets=[cfg5["sheets"][k]], hcdefs=cfg5["hcdefs"], elevmask=[cfg5["elevmask"][k]])


@pytest.mark.parametrize("scale,correctA", [(True, True), (False, False)])
def test_greenland_1km_six_matrices_against_oracle(cfg5, scale, correctA):
    cfg = one_sheet(cfg5, 0)
    sh, em = cfg["sheets"][0], cfg["elevmask"][0]
    assert (sh.nx, sh.ny) == (1501, 2801)
    go, gg = make_oracle(cfg), make_gpu(cfg)
    rm = gg.regrid_matrices(sh.name, em, scale=scale, correctA=correctA)
    for spec in ("EvI", "AvI", "IvE", "IvA", "EvA", "AvE"):
        Wo = go.matrix(0, em, spec, scale=scale, correctA=correctA)
        Wg = rm.matrix(spec)
        assert Wg.shape == Wo.shape and Wg.nnz == Wo.nnz, spec
        assert np.array_equal(Wg.dim(0), Wo.dims[0]) and np.array_equal(Wg.dim(1), Wo.dims[1]), spec + ": dense<->sparse maps differ"
        r, c, v = sorted_coo(Wg)
        assert np.array_equal(r, Wo.row) and np.array_equal(c, Wo.col), spec + ": sparsity pattern differs"
        np.testing.assert_allclose(v, Wo.val, rtol=1e-12, atol=0, err_msg=spec)
        np.testing.assert_allclose(Wg.get_weights(0), Wo.wM, rtol=1e-12, atol=0, err_msg=spec + " wM")
        np.testing.assert_allclose(Wg.get_weights(1), Wo.Mw, rtol=1e-12, atol=0, err_msg=spec + " Mw")


def test_config5_two_sheets_merged_against_oracle(cfg5):
    """Greenland 1 km + Antarctica 1 km (36 M ice cells) on one GCMRegridder.  GPU: per sheet the unscaled EvI and AvI
    (the coupler's E1vI_unscaled / A1vI_unscaled, IceCoupler.cpp:371-377) applied in one pass to 4 fields, contributions
    merged with the library's VectorMultivec twin.  Oracle: its own matrices, its own apply, orc.multivec_to_dense."""
    gA, hc = cfg5["gridA"], cfg5["hcdefs"]
    go, gg = make_oracle(cfg5), make_gpu(cfg5)
    nvar = 4
    rng = np.random.default_rng(11)
    contrib = {"EvI": [], "AvI": []}
    for k, (sh, em) in enumerate(zip(cfg5["sheets"], cfg5["elevmask"])):
        rm = gg.regrid_matrices(sh.name, em, scale=False, correctA=True)
        Wg = {s: rm.matrix(s) for s in ("EvI", "AvI")}
        Wo = {s: go.matrix(k, em, s, scale=False, correctA=True) for s in ("EvI", "AvI")}
        for s in ("EvI", "AvI"):
            assert np.array_equal(Wg[s].dim(0), Wo[s].dims[0]) and np.array_equal(Wg[s].dim(1), Wo[s].dims[1]), (sh.name, s)
            r, c, v = sorted_coo(Wg[s])
            assert np.array_equal(r, Wo[s].row) and np.array_equal(c, Wo[s].col), (sh.name, s)
            np.testing.assert_allclose(v, Wo[s].val, rtol=1e-12, atol=0)
        X = rng.uniform(0.0, 1.0, (Wo["EvI"].shape[1], nvar))
        YE, YA = api.apply_many([Wg["EvI"], Wg["AvI"]], X, layout="cell")
        for s, Y in (("EvI", YE), ("AvI", YA)):
            Yo = Wo[s].apply(np.ascontiguousarray(X.T)).T
            A = abs_rowsum(Wo[s], X)
            assert np.all(np.abs(Y - Yo) <= 1e-12 * A + 1e-300), (sh.name, s)
            contrib[s].append((Wo[s].dims[0], Wg[s].get_weights(0), Y, Wo[s].wM, Yo))
        del Wg, Wo, rm
    for s, n_sparse in (("EvI", gg.nE), ("AvI", gg.nA)):
        idx = np.concatenate([c[0] for c in contrib[s]])
        scale_g, dense_g = api.multivec_to_dense(idx, np.concatenate([c[1] for c in contrib[s]]), np.concatenate([c[2] for c in contrib[s]]), n_sparse)
        scale_o, dense_o = orc.multivec_to_dense(idx, np.concatenate([c[3] for c in contrib[s]]), np.concatenate([c[4] for c in contrib[s]]), n_sparse)
        touched = np.zeros(n_sparse, bool); touched[idx] = True
        assert np.array_equal(np.isnan(dense_g), np.isnan(dense_o))
        np.testing.assert_allclose(scale_g[touched], scale_o[touched], rtol=1e-12)
        np.testing.assert_allclose(dense_g[:, touched], dense_o[:, touched], rtol=1e-11, atol=1e-13)
        # the merged field of a constant is that constant wherever any sheet contributes (multivec.cpp:35-81)
    ones = [np.ones((c[0].shape[0], 1)) for c in contrib["AvI"]]
    del ones


def abs_rowsum(Wo, X):
    import scipy.sparse as sp
    A = sp.csr_matrix((np.abs(Wo.val), (Wo.row, Wo.col)), shape=Wo.shape)
    return A @ np.abs(X)
