"""GPU parity of the matrix build: CUDA path (through the C ABI) vs the CPU oracle on identical synthetic grids.
Gates (BASELINE.json north_star): sparsity pattern and dense<->sparse maps bit-exact; values <= 1e-12 relative."""
import itertools

import numpy as np
import pytest

from conftest import make_gpu, make_oracle

pytestmark = pytest.mark.gpu

SIX = ["AvI", "IvA", "EvI", "IvE", "EvA", "AvE"]
XS = ["AvX", "XvA", "EvX", "XvE"]
RTOL = 1e-12


def csc_order(row, col, val):
    o = np.lexsort((row, col))
    return row[o], col[o], val[o]


def assert_same(Wg, Wo, tag):
    assert Wg.shape == Wo.shape, tag
    assert Wg.nnz == Wo.nnz, tag
    assert Wg.conservative == Wo.conservative, tag
    for k in (0, 1):
        assert np.array_equal(Wg.dim(k), Wo.dims[k]), "%s: dims[%d] differ" % (tag, k)
    r, c, v = csc_order(*Wg.to_coo_dense())
    assert np.array_equal(r, Wo.row) and np.array_equal(c, Wo.col), tag + ": pattern differs"
    np.testing.assert_allclose(v, Wo.val, rtol=RTOL, atol=0, err_msg=tag)
    np.testing.assert_allclose(Wg.get_weights(0), Wo.wM, rtol=RTOL, atol=0, err_msg=tag + " wM")
    np.testing.assert_allclose(Wg.get_weights(1), Wo.Mw, rtol=RTOL, atol=0, err_msg=tag + " Mw")
    return np.array_equal(v, Wo.val) and np.array_equal(Wg.get_weights(0), Wo.wM) and np.array_equal(Wg.get_weights(1), Wo.Mw)


@pytest.mark.parametrize("spec", SIX + XS)
def test_cfg1_all_specs(cfg1, spec):
    go, gg = make_oracle(cfg1), make_gpu(cfg1)
    em = cfg1["elevmask"][0]
    exact = True
    for scale, cA in itertools.product((True, False), (True, False)):
        Wo = go.matrix(0, em, spec, scale=scale, correctA=cA)
        Wg = gg.regrid_matrices(cfg1["sheets"][0].name, em, scale=scale, correctA=cA).matrix(spec)
        exact &= assert_same(Wg, Wo, "%s scale=%d correctA=%d" % (spec, scale, cA))
    # stronger than the gate: the build keeps the oracle's operation order, so values are bit-identical
    assert exact, spec + ": values agree to 1e-12 but are not bit-identical"


@pytest.mark.parametrize("spec", SIX)
def test_cfg2_six(cfg2, spec):
    go, gg = make_oracle(cfg2), make_gpu(cfg2)
    em = cfg2["elevmask"][0]
    for scale, cA in ((True, True), (False, False)):
        Wo = go.matrix(0, em, spec, scale=scale, correctA=cA)
        Wg = gg.regrid_matrices(cfg2["sheets"][0].name, em, scale=scale, correctA=cA).matrix(spec)
        assert_same(Wg, Wo, "%s scale=%d correctA=%d" % (spec, scale, cA))


def test_elev_class_interp(cfg1):
    go, gg = make_oracle(cfg1, 1), make_gpu(cfg1, "ELEV_CLASS_INTERP")
    em = cfg1["elevmask"][0]
    for spec in ("EvI", "IvE", "EvA", "AvE"):
        Wo = go.matrix(0, em, spec, scale=True, correctA=True)
        Wg = gg.regrid_matrices(cfg1["sheets"][0].name, em, scale=True, correctA=True).matrix(spec)
        assert_same(Wg, Wo, spec)


def test_prepopulated_dims(cfg1):
    """matrix_d with caller dims (global_ec.cpp:522-627 shares dims; IceCoupler.cpp:366 passes an identity dimI)."""
    from icebin_b200 import api
    go, gg = make_oracle(cfg1), make_gpu(cfg1)
    sh = cfg1["sheets"][0]
    em = cfg1["elevmask"][0]
    rm = gg.regrid_matrices(sh.name, em)
    # (a) identity dimI of all nI cells, dimE pre-sized empty
    dimE, dimI = api.SparseSet(), api.SparseSet(identity=sh.nI)
    Wg = rm.matrix_d("EvI", (dimE, dimI), scale=False, correctA=False)
    Wo = go.matrix(0, em, "EvI", scale=False, correctA=False, dims=(np.zeros(0, np.int64), np.arange(sh.nI)))
    assert_same(Wg, Wo, "EvI identity dimI")
    # (b) the same dims re-used for the next matrix: ids are kept, nothing new appended
    nE0 = dimE.dense_extent
    Wg2 = rm.matrix_d("IvE", (dimI, dimE), scale=True, correctA=True)
    Wo2 = go.matrix(0, em, "IvE", scale=True, correctA=True, dims=(np.arange(sh.nI), Wo.dims[0]))
    assert dimE.dense_extent == nE0
    assert_same(Wg2, Wo2, "IvE shared dims")
    # (c) partially pre-populated, shuffled A ids (some never touched => empty rows, weight 0)
    pre = cfg1["gridA"].dim[::3][::-1].copy()
    dA, dI = api.SparseSet(pre, sparse_extent=cfg1["gridA"].nA), api.SparseSet()
    Wg3 = rm.matrix_d("AvI", (dA, dI), scale=True, correctA=True)
    Wo3 = go.matrix(0, em, "AvI", scale=True, correctA=True, dims=(pre, None))
    assert_same(Wg3, Wo3, "AvI prepopulated")


def test_errors(cfg1):
    from icebin_b200 import api
    gg = make_gpu(cfg1)
    sh = cfg1["sheets"][0]
    em = cfg1["elevmask"][0].copy()
    rm = gg.regrid_matrices(sh.name, em)
    with pytest.raises(KeyError):
        rm.matrix("IvI")
    em[np.flatnonzero(~np.isnan(em))[0]] = 1e5      # above the top elevation class: fatal upstream (L0.cpp:84-85)
    with pytest.raises(api.IceBinError) as ei:
        gg.regrid_matrices(sh.name, em).matrix("EvI")
    assert ei.value.code == -3 and "out of bounds" in str(ei.value)
    gg.regrid_matrices(sh.name, em).matrix("AvI")   # A matrices never look at hcdefs


def test_all_masked(cfg1):
    gg, go = make_gpu(cfg1), make_oracle(cfg1)
    sh = cfg1["sheets"][0]
    em = np.full(sh.nI, np.nan)
    W = gg.regrid_matrices(sh.name, em).matrix("EvI")
    assert W.shape == (0, 0) and W.nnz == 0


def _shuffled_with_duplicates(cfg, seed=3):
    """The same grids with the exchange cells in RANDOM order (upstream's order is implementation-defined,
    SURVEY Appendix B.9) and 5 % of them split into two records of half the area (duplicate (iA, iI) pairs, which
    setFromTriplets sums)."""
    import copy
    rng = np.random.default_rng(seed)
    c = dict(cfg)
    sh = copy.copy(cfg["sheets"][0])
    n = sh.nX
    dup = rng.choice(n, n // 20, replace=False)
    ij = np.concatenate([sh.ex_ij, sh.ex_ij[dup]])
    area = sh.ex_area.copy()
    area[dup] *= 0.5
    area = np.concatenate([area, area[dup]])
    perm = rng.permutation(len(area))
    sh.ex_ij, sh.ex_area = np.ascontiguousarray(ij[perm]), np.ascontiguousarray(area[perm])
    c["sheets"] = [sh]
    return c


@pytest.mark.parametrize("spec", SIX + ["EvX", "XvA"])
def test_arbitrary_order_and_duplicate_cells(cfg1, spec):
    cfg = _shuffled_with_duplicates(cfg1)
    go, gg = make_oracle(cfg), make_gpu(cfg)
    em = cfg["elevmask"][0]
    for scale, cA in ((True, True), (False, False)):
        Wo = go.matrix(0, em, spec, scale=scale, correctA=cA)
        Wg = gg.regrid_matrices(cfg["sheets"][0].name, em, scale=scale, correctA=cA).matrix(spec)
        exact = assert_same(Wg, Wo, "%s shuffled scale=%d correctA=%d" % (spec, scale, cA))
        assert exact, spec + ": duplicates must be summed in emission order (stable sort) to stay bit-identical"
    # and the result is the same operator as with the tidy grid, up to the dense renumbering
    if spec in SIX:
        W0 = make_oracle(cfg1).matrix(0, em, spec, scale=True, correctA=True)
        W1 = go.matrix(0, em, spec, scale=True, correctA=True)
        import scipy.sparse as sp
        A0 = sp.coo_matrix((W0.val, (W0.dims[0][W0.row], W0.dims[1][W0.col])), shape=W0.sparse_extent).tocsr()
        A1 = sp.coo_matrix((W1.val, (W1.dims[0][W1.row], W1.dims[1][W1.col])), shape=W1.sparse_extent).tocsr()
        assert abs(A0 - A1).max() <= 1e-12 * abs(A0).max()


def test_single_elevation_class_is_fatal(cfg1):
    """linterp_1d_b with one height point always runs off the top (IceRegridder_L0.cpp:82-85)."""
    from icebin_b200 import api
    c = dict(cfg1); c["hcdefs"] = np.array([0.0])
    gg = make_gpu(c)
    with pytest.raises(api.IceBinError) as e:
        gg.regrid_matrices(c["sheets"][0].name, c["elevmask"][0]).matrix("EvI")
    assert e.value.code == -3
    gg.regrid_matrices(c["sheets"][0].name, c["elevmask"][0]).matrix("IvA")


def test_elevation_on_class_boundaries(cfg1):
    """Elevations exactly on a height point give one zero weight, which upstream drops (whps != 0, L0.cpp:133,139)."""
    go, gg = make_oracle(cfg1), make_gpu(cfg1)
    em = cfg1["elevmask"][0].copy()
    ok = np.flatnonzero(~np.isnan(em))
    em[ok[::3]] = cfg1["hcdefs"][3]          # exactly on a point
    em[ok[1::3]] = -25.0                     # negative: clamped to 0 (L0.cpp:123)
    for spec in ("EvI", "IvE", "EvA", "AvE"):
        Wo = go.matrix(0, em, spec, scale=True, correctA=True)
        Wg = gg.regrid_matrices(cfg1["sheets"][0].name, em, scale=True, correctA=True).matrix(spec)
        assert assert_same(Wg, Wo, spec)
