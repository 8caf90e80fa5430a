"""CPU: the scipy restatement of GCMRegridder_ModelE (oracle/oracle_np.py:ModelE) against properties the reference states
(GCMRegridder_ModelE.cpp:275-282 "should consist only of 1's and 0's"; :354-375 "M_unscaled = wM * M_scaled"; weights of
an atmosphere cell are the sum over its ocean-grid cells) and against a dense evaluation of the same formulas."""
import numpy as np

from conftest import make_oracle
from modele_common import SPEC_O, case
from oracle import oracle_np


def setup():
    cfg, fp, fm = case()
    go = make_oracle(cfg)
    sh = cfg["sheets"][0]
    return cfg, fp, fm, oracle_np.ModelE(go, 0, cfg["elevmask"][0], fp, fm, SPEC_O, 6371000.0, len(cfg["hcdefs"]), sh.nI)


def test_helper_invariants():
    cfg, fp, fm, me = setup()
    for X in ("A", "E"):
        H = me.helper(X, "I")
        T = H["XAmvXOm"]
        Z = (T @ __import__("scipy.sparse").sparse.diags(np.where(np.isfinite(H["XAmvXOms"]), H["XAmvXOms"], 0.0))).tocoo()
        assert np.allclose(Z.data, 1.0, rtol=1e-14)                    # :275-276
        assert np.all(np.diff(T.tocsc().indptr) <= 1)                  # every ocean-grid cell lies in one atmosphere cell (zero-weight ECs: none)
        assert set(H["dimXOm"]) <= set(H["dimXOp"])                    # ModelE's continent is a subset of the ice model's cells
        assert all(fm[s % me.nO] == 0 for s in H["dimXOm"])
        assert np.all(H["wXAm"] >= 0) and H["wXAm"].sum() > 0


def test_scaled_and_unscaled_agree_and_rows_sum_to_one():
    cfg, fp, fm, me = setup()
    for X in ("A", "E"):
        d0, d1, Ms, wM, Mw = me.XAmvGp(X, "I", scale=True)
        _, _, Mu, wMu, _ = me.XAmvGp(X, "I", scale=False)
        assert Ms.shape == (len(d0), len(d1)) and np.array_equal(wM, wMu)
        D = (Mu - __import__("scipy.sparse").sparse.diags(wM) @ Ms).tocoo()
        assert np.abs(D.data).max(initial=0) <= 1e-12 * np.abs(Mu.data).max()          # "M_unscaled = wM * M_scaled"
        rs = np.asarray(Ms.sum(1)).ravel()
        live = wM > 0
        assert np.allclose(rs[live], 1.0, rtol=1e-12)                  # a constant ice field regrids to the same constant
        # and the way back
        i0, x1, Mb, wMb, Mwb = me.GpvXAm("I", X, scale=True)
        assert np.array_equal(i0, d1) and np.array_equal(x1, d0) and np.allclose(Mwb, wM, rtol=1e-14)
        rb = np.asarray(Mb.sum(1)).ravel()
        assert np.all((np.abs(rb - 1.0) < 1e-12) | (rb == 0.0) | (rb < 1.0 + 1e-12))
