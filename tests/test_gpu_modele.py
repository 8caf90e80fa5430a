"""GPU parity of GCMRegridder_ModelE (icebin_b200/modele.py + csrc/modele.cu) against the scipy restatement built on the C++
oracle's ocean-grid matrices (oracle/oracle_np.py:ModelE): dense<->sparse maps and sparsity pattern bit for bit, values and
weights to 1e-12, for the eight ice / exchange <-> atmosphere / elevation matrices, scaled and unscaled."""
import numpy as np
import pytest
import scipy.sparse as sp

from conftest import make_gpu, make_oracle
from icebin_b200 import api, modele
from modele_common import SPEC_O, case
from oracle import oracle_np

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def both():
    cfg, fp, fm = case()
    sh = cfg["sheets"][0]
    go, gg = make_oracle(cfg), make_gpu(cfg)
    gA = modele.GCMRegridder_ModelE(gg, api.HntrSpec(*SPEC_O), 6371000.0)
    assert (gA.hspecA.im, gA.hspecA.jm) == (144, 90) and gA.nA == 144 * 90
    return cfg, fp, fm, go, gA


@pytest.mark.parametrize("scale", [True, False])
@pytest.mark.parametrize("spec", ["AvI", "EvI", "IvA", "IvE", "AvX", "EvX", "XvA", "XvE"])
def test_matrices_against_oracle(both, spec, scale):
    cfg, fp, fm, go, gA = both
    sh, em = cfg["sheets"][0], cfg["elevmask"][0]
    G = "I" if "I" in spec else "X"
    X = "A" if "A" in spec else "E"
    me = oracle_np.ModelE(go, 0, em, fp, fm, SPEC_O, 6371000.0, len(cfg["hcdefs"]), sh.nI if G == "I" else sh.nX)
    d0, d1, Mo, wM, Mw = me.XAmvGp(X, G, scale) if spec[0] in "AE" else me.GpvXAm(G, X, scale)
    Wg = gA.regrid_matrices(sh.name, fp, fm, em, scale=scale).matrix(spec)
    assert Wg.shape == Mo.shape and not Wg.conservative
    assert np.array_equal(Wg.dim(0), d0) and np.array_equal(Wg.dim(1), d1), "dense<->sparse maps differ"
    r, c, v = Wg.to_coo_dense()
    Mg = sp.csr_matrix((v, (r, c)), shape=Wg.shape)
    Mo.sum_duplicates(); Mo.sort_indices(); Mg.sort_indices()
    assert np.array_equal(Mg.indptr, Mo.indptr) and np.array_equal(Mg.indices, Mo.indices), "sparsity pattern differs"
    np.testing.assert_allclose(Mg.data, Mo.data, rtol=1e-12, atol=0)
    np.testing.assert_allclose(Wg.get_weights(0), wM, rtol=1e-12, atol=0)
    np.testing.assert_allclose(Wg.get_weights(1), Mw, rtol=1e-12, atol=0)
    assert Wg.sparse_extent == ((gA.nAE(X), me.nG) if spec[0] in "AE" else (me.nG, gA.nAE(X)))


def test_apply_and_aliases(both):
    """AAmvIp is AvI (:524-527); a constant ice field regrids to the same constant on every atmosphere cell ModelE keeps;
    AOmvAAm's rows add up to the ocean-grid cell areas (Hntr overlap, tests/test_hntr.cpp:332-399)."""
    cfg, fp, fm, go, gA = both
    sh, em = cfg["sheets"][0], cfg["elevmask"][0]
    rm = gA.regrid_matrices(sh.name, fp, fm, em)
    A1, A2 = rm.matrix("AvI"), rm.matrix("AAmvIp")
    assert np.array_equal(A1.to_csr()[2], A2.to_csr()[2]) and np.array_equal(A1.dim(0), A2.dim(0))
    one = np.ones((1, A1.shape[1]))
    y = A1.apply(one, force_conservation=False)
    live = A1.get_weights(0) > 0
    assert np.allclose(y[0, live], 1.0, rtol=1e-12)
    with pytest.raises(KeyError):
        rm.matrix("AvE")
    O = rm.matrix("AOmvAAm")
    T = rm.matrix("AAmvAOm")
    assert O.shape == T.shape[::-1] and np.array_equal(O.dim(0), T.dim(1))
    dxyp = api.Hntr(17.17, api.HntrSpec(*SPEC_O), gA.hspecA).get("dxypB")
    jO = O.dim(0) // SPEC_O[0]
    np.testing.assert_allclose(O.get_weights(0), 6371000.0 ** 2 * dxyp[jO], rtol=1e-12)
