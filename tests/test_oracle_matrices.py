"""CPU: pins the oracle (oracle/oracle.cpp).

  * against the independent scipy twin (oracle/oracle_np.py) -- pattern and dense-index maps bit-equal, values 1e-13;
  * against every assertion of the reference's own demo test tests/test_conserv/test_conserv.py:89-205
    (the reference holds no golden vectors for the six matrices: 'parity unpinned', properties only);
  * against the doc example of IceRegridder.hpp:150-155 and the committed regression vectors in tests/golden/.
The mask used for the property tests is the reference's own fixture elev_mask.cdl (SeaRISE 20 km, 76x141)."""
import os

import numpy as np
import pytest

from conftest import make_oracle
from icebin_b200 import synth
from oracle import orc, oracle_np

GOLD = os.path.join(os.path.dirname(__file__), "golden")
SIX = ["AvI", "IvA", "EvI", "IvE", "EvA", "AvE"]


@pytest.fixture(scope="module")
def searise20():
    """ModelE 2x2.5 + SeaRISE-shaped 20 km grid with the reference's elev_mask.cdl as elevmaskI (emI_ice)."""
    d = np.load(os.path.join(GOLD, "elevmask_searise20km.npz"))
    gA, sh = synth.make_gcm_and_sheet("greenland_20km")
    assert (sh.nx, sh.ny) == (int(d["nx"]), int(d["ny"]))
    return dict(gridA=gA, sheets=[sh], hcdefs=synth.hcdefs_conserv(40), elevmask=[d["emI_ice"]], emI_land=d["emI_land"])


def test_linterp_doc_example():
    idx, w = orc.linterp([3, 5, 6, 8], 6.5)      # IceRegridder.hpp:150-155
    assert list(idx) == [2, 3] and np.allclose(w, [.75, .25])
    idx, w = orc.linterp([3, 5, 6, 8], 1.0)      # below the bottom: extrapolates with i1 = 1
    assert list(idx) == [0, 1] and np.allclose(w, [2.0, -1.0])
    idx, w = orc.linterp([3, 5, 6, 8], 5.0)      # exactly on a point: lower_bound puts it at the TOP of (3,5]
    assert list(idx) == [0, 1] and np.allclose(w, [0.0, 1.0])
    with pytest.raises(orc.OracleError):
        orc.linterp([3, 5, 6, 8], 8.5)           # IceRegridder_L0.cpp:84-85
    assert orc.nearest([3, 5, 6, 8], 4.0) == 0   # ties go down (d0 <= d1)
    assert orc.nearest([3, 5, 6, 8], 100.0) == 3 and orc.nearest([3, 5, 6, 8], -5) == 0


@pytest.mark.parametrize("correctA", [False, True])
@pytest.mark.parametrize("scale", [False, True])
def test_oracle_vs_scipy_twin(cfg1, correctA, scale):
    g = make_oracle(cfg1)
    gA, sh, hc, em = cfg1["gridA"], cfg1["sheets"][0], cfg1["hcdefs"], cfg1["elevmask"][0]
    for spec in SIX:
        W = g.matrix(0, em, spec, scale=scale, correctA=correctA)
        T = oracle_np.matrix(gA, sh, hc, em, spec, scale=scale, correctA=correctA)
        assert np.array_equal(W.dims[0], T["dims"][0]) and np.array_equal(W.dims[1], T["dims"][1]), spec
        M = T["M"].tocoo()
        o = np.lexsort((M.row, M.col))
        assert np.array_equal(M.row[o], W.row) and np.array_equal(M.col[o], W.col), spec
        np.testing.assert_allclose(M.data[o], W.val, rtol=1e-13, err_msg=spec)
        np.testing.assert_allclose(T["wM"], W.wM, rtol=1e-13, err_msg=spec)
        np.testing.assert_allclose(T["Mw"], W.Mw, rtol=1e-13, err_msg=spec)


def test_sizes_match_survey(cfg1):
    """SURVEY.md Appendix F sizes for config 1."""
    g = make_oracle(cfg1)
    em = cfg1["elevmask"][0]
    sh = cfg1["sheets"][0]
    assert sh.nX == 3290 and int(np.sum(~np.isnan(em[sh.ex_ij[:, 1]]))) == 1362
    shapes = {s: (g.matrix(0, em, s).shape, g.matrix(0, em, s).nnz) for s in ("AvI", "EvI", "EvA")}
    assert shapes == {"AvI": ((110, 727), 1362), "EvI": ((442, 727), 2724), "EvA": ((442, 110), 442)}


def test_golden_regression(cfg1):
    gold = np.load(os.path.join(GOLD, "oracle_golden_cfg1.npz"))
    g = make_oracle(cfg1)
    for spec in SIX:
        W = g.matrix(0, cfg1["elevmask"][0], spec, scale=True, correctA=True)
        assert np.array_equal(W.row, gold[spec + "_row"]) and np.array_equal(W.col, gold[spec + "_col"])
        assert np.array_equal(W.dims[0], gold[spec + "_d0"]) and np.array_equal(W.dims[1], gold[spec + "_d1"])
        np.testing.assert_allclose(W.val, gold[spec + "_val"], rtol=1e-14)
        np.testing.assert_allclose(W.wM, gold[spec + "_wM"], rtol=1e-14)
        np.testing.assert_allclose(W.Mw, gold[spec + "_Mw"], rtol=1e-14)


@pytest.mark.parametrize("correctA", [False, True])
def test_conserv_properties(searise20, correctA):
    """test_conserv.py: (89-111) unscaled == diag(w) scaled; (132-172) constant round trips; (180-205) mass conservation."""
    cfg = searise20
    g = make_oracle(cfg)
    em = cfg["elevmask"][0]
    sh = cfg["sheets"][0]
    S = {s: g.matrix(0, em, s, scale=True, correctA=correctA) for s in SIX}
    U = {s: g.matrix(0, em, s, scale=False, correctA=correctA) for s in SIX}
    for s in SIX:
        assert np.array_equal(S[s].row, U[s].row) and np.array_equal(S[s].col, U[s].col)
        np.testing.assert_allclose(U[s].val, S[s].wM[S[s].row] * S[s].val, rtol=1e-13, err_msg=s)
    # constant round trips
    for a, b in (("AvI", "IvA"), ("EvI", "IvE"), ("AvE", "EvA"), ("EvA", "AvE")):
        assert np.array_equal(S[a].dims[0], S[b].dims[1]) and np.array_equal(S[a].dims[1], S[b].dims[0])
        one = np.ones((1, S[a].shape[1]))
        assert np.abs(S[b].apply(S[a].apply(one)) - 1).max() < 1e-12, (a, b)
    # A <- I of a constant is NOT the constant when projection areas are corrected (test_conserv.py:143-149)
    if correctA:
        assert np.abs(S["AvI"].apply(np.ones((1, S["AvI"].shape[1]))) - 1).max() > 1e-3
    # mass conservation for the 'diagonal' field i+j and the elevation field
    ii, jj = np.divmod(np.arange(sh.nI), sh.ny)
    for fI in ((ii + jj).astype(float), np.nan_to_num(em)):
        f = fI[S["AvI"].dims[1]][None, :]
        mass = (f[0] * S["AvI"].Mw).sum()
        for s in ("AvI", "EvI"):
            assert abs((S[s].apply(f)[0] * S[s].wM).sum() / mass - 1) < 1e-12, s
        for a, b in (("AvI", "EvA"), ("EvI", "AvE")):       # I -> A -> E  and I -> E -> A keep the integral
            y = S[b].apply(S[a].apply(f))
            assert abs((y[0] * S[b].wM).sum() / mass - 1) < 1e-12, (a, b)
    # EvA is 'local': iE mod nA == iA (combine_global_ec.cpp:229-245) and has one entry per E row
    W = S["EvA"]
    assert W.nnz == W.shape[0]
    assert np.array_equal(W.dims[0][W.row] % cfg["gridA"].nA, W.dims[1][W.col])


def test_weights_are_areas(searise20):
    """wM of AvI (no correctA) is the ice-covered area of each A cell; Mw is the covered area of each I cell."""
    cfg = searise20
    g = make_oracle(cfg)
    em, sh = cfg["elevmask"][0], cfg["sheets"][0]
    W = g.matrix(0, em, "AvI", scale=True, correctA=False)
    ok = ~np.isnan(em[sh.ex_ij[:, 1]])
    areaA = np.bincount(sh.ex_ij[ok, 0], sh.ex_area[ok], cfg["gridA"].nA)
    areaI = np.bincount(sh.ex_ij[ok, 1], sh.ex_area[ok], sh.nI)
    np.testing.assert_allclose(W.wM, areaA[W.dims[0]], rtol=1e-12)
    np.testing.assert_allclose(W.Mw, areaI[W.dims[1]], rtol=1e-12)
    np.testing.assert_allclose(W.Mw, sh.dx * sh.dx, rtol=1e-12)      # every unmasked ice cell is fully covered


def test_land_mask_is_superset(searise20):
    cfg = searise20
    g = make_oracle(cfg)
    Wi = g.matrix(0, cfg["elevmask"][0], "IvA")
    Wl = g.matrix(0, cfg["emI_land"], "IvA")
    assert set(Wi.dims[0]) < set(Wl.dims[0])


def test_prepopulated_dims_keep_ids(cfg1):
    g = make_oracle(cfg1)
    em = cfg1["elevmask"][0]
    pre = cfg1["gridA"].dim[::3][::-1].copy()
    W = g.matrix(0, em, "AvI", dims=(pre, None))
    assert np.array_equal(W.dims[0][: len(pre)], pre)
    fresh = g.matrix(0, em, "AvI")
    assert set(W.dims[0]) == set(pre) | set(fresh.dims[0])
    assert np.all(W.wM[: len(pre)][~np.isin(pre, fresh.dims[0])] == 0)      # untouched rows: weight 0, no entries


def test_unknown_spec_and_errors(cfg1):
    g = make_oracle(cfg1)
    em = cfg1["elevmask"][0].copy()
    with pytest.raises(KeyError):
        g.matrix(0, em, "IvI")
    em[np.flatnonzero(~np.isnan(em))[0]] = 1e5
    with pytest.raises(orc.OracleError):
        g.matrix(0, em, "EvI")


def test_smoothing_oracle(cfg1):
    """smoother.cpp: rows of smoothI sum to 1; the smoothed IvE is flagged non-conservative and apply() restores the integral."""
    g = make_oracle(cfg1)
    em = cfg1["elevmask"][0]
    sig = (100e3, 100e3, 300.0)
    W0 = g.matrix(0, em, "IvE", scale=True, correctA=True)
    Ws = g.matrix(0, em, "IvE", scale=True, correctA=True, sigma=sig)
    assert W0.conservative and not Ws.conservative and Ws.nnz > W0.nnz
    Wn = g.matrix(0, em, "IvE", scale=True, correctA=False, sigma=sig)     # without the projection ratio rows sum to 1
    assert np.abs(Wn.apply(np.ones((1, Wn.shape[1])), force_conservation=False) - 1).max() < 1e-12
    x = np.random.default_rng(0).uniform(0, 1, (3, Ws.shape[1]))
    y = Ws.apply(x, force_conservation=True)
    np.testing.assert_allclose((y * Ws.wM).sum(1), (x * Ws.Mw).sum(1), rtol=1e-12)
    y0 = Ws.apply(x, force_conservation=False)
    assert np.abs((y0 * Ws.wM).sum(1) / (x * Ws.Mw).sum(1) - 1).max() > 1e-6


def test_multivec_merge():
    """multivec.cpp:35-81: out = sum(val)/sum(weight) per index; untouched -> fill."""
    idx = np.array([3, 1, 3, 0]); w = np.array([2.0, 1.0, 6.0, 4.0]); v = np.array([[4.0, 1], [5.0, 2], [12.0, 3], [8.0, 4]])
    scale, dense = orc.multivec_to_dense(idx, w, v, 5, fill=-1.0)
    assert np.allclose(scale[[0, 1, 3]], [0.25, 1.0, 0.125]) and np.isinf(scale[2])
    assert np.allclose(dense[0], [2.0, 5.0, -1.0, 2.0, -1.0]) and np.allclose(dense[1], [1.0, 2.0, -1.0, 0.5, -1.0])
