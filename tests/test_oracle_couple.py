"""CPU: the oracle's restatement of the coupling step's regridding side (variable transformation, E1vE0c) against
independent numpy/scipy statements of the same algebra (e1ve0.cpp:55-106, IceCoupler.cpp:203-252)."""
import numpy as np
import scipy.sparse as sp

from conftest import make_oracle
from couple_common import stepped_elevmask, two_sheet_config, var_transform
from oracle import orc


def test_vartrans_matches_numpy_and_skips_structural_zeros():
    M, b = var_transform(6, 9, 1)
    O = np.random.default_rng(0).uniform(-3, 3, (50, 6))
    Z = orc.vartrans(O, M, b)
    assert np.allclose(Z, O @ M.toarray() + b, rtol=1e-14, atol=1e-14)
    # a NaN input variable only reaches the outputs that name it (contracts::ALLOW_NAN)
    O2 = O.copy(); O2[:, 2] = np.nan
    Z2 = orc.vartrans(O2, M, b)
    uses2 = np.asarray((M.toarray()[2] != 0))
    assert np.isnan(Z2[:, uses2]).all() and np.array_equal(Z2[:, ~uses2], Z[:, ~uses2])


def scipy_e1ve0c(XuE1s, XuE0s, nE):
    tot = None
    sE1 = np.zeros(nE)
    for W1, W0 in zip(XuE1s, XuE0s):
        def sparsified(W):
            return sp.csr_matrix((W.val, (W.row, W.dims[1][W.col])), shape=(W.shape[0], nE))
        X1, X0 = sparsified(W1), sparsified(W0)
        with np.errstate(divide="ignore"):
            D = sp.diags(np.where(W0.wM != 0, 1 / W0.wM, 0)) @ X0 - sp.diags(np.where(W1.wM != 0, 1 / W1.wM, 0)) @ X1
        loc = (X1.T @ D).tocsr()
        tot = loc if tot is None else tot + loc
        np.add.at(sE1, W1.dims[1], W1.Mw)
    with np.errstate(divide="ignore"):
        return (sp.diags(np.where(sE1 != 0, 1 / sE1, 0)) @ tot).tocoo(), sE1


def xue(go, sheet, em):
    nX = None
    W = go.matrix(sheet, em, "XvE", scale=False, correctA=True, dims=(np.arange(go_nX(go, sheet)), None))
    return W


def go_nX(go, sheet):
    return go._nX[sheet]


def make(cfg):
    go = make_oracle(cfg)
    go._nX = [sh.nX for sh in cfg["sheets"]]
    return go


def test_e1ve0c_two_sheets_matches_scipy():
    cfg = two_sheet_config()
    go = make(cfg)
    nE = cfg["gridA"].nA * len(cfg["hcdefs"])
    em0 = cfg["elevmask"]
    em1 = [stepped_elevmask(em0[i], 10 + i) for i in range(2)]
    X0 = [xue(go, i, em0[i]) for i in range(2)]
    X1 = [xue(go, i, em1[i]) for i in range(2)]
    assert all(W.shape[0] == cfg["sheets"][i].nX for i, W in enumerate(X0))
    r, c, v = orc.e1ve0c(X1, X0, nE)
    assert np.all(np.diff(r * nE + c) > 0)                      # sorted, consolidated
    ref, sE1 = scipy_e1ve0c(X1, X0, nE)
    got = sp.coo_matrix((v, (r, c)), shape=(nE, nE))
    diff = abs(got - ref)
    scale = abs(ref).max()
    assert diff.max() <= 1e-12 * scale
    # the sheets do share elevation-class cells (the case consolidate() exists for)
    shared = np.intersect1d(X1[0].dims[1], X1[1].dims[1])
    assert len(shared) > 0
    # rows of the correction only name E1 cells, columns E0 or E1 cells
    e1 = np.union1d(X1[0].dims[1], X1[1].dims[1]); e0 = np.union1d(X0[0].dims[1], X0[1].dims[1])
    assert np.isin(r, e1).all() and np.isin(c, np.union1d(e0, e1)).all()


def test_e1ve0c_is_zero_when_nothing_changed(cfg1):
    go = make(cfg1)
    nE = cfg1["gridA"].nA * len(cfg1["hcdefs"])
    X = [xue(go, 0, cfg1["elevmask"][0])]
    r, c, v = orc.e1ve0c(X, X, nE)
    assert len(v) > 0 and np.all(v == 0.0)
