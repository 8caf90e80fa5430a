"""GPU parity of the one-pass apply for wide matrices (apply_stream.cu): EvI alone and the stacked [EvI; AvI] of the
coupling step (IceCoupler.cpp:384-458), against the oracle's Eigen-order products and against scipy for generic
matrices.  Tolerance: |dY| <= 1e-12 * sum|terms| of the row (the north star's 1e-12 relative gate)."""
import numpy as np
import pytest
import scipy.sparse as sp

from conftest import make_gpu, make_oracle

pytestmark = pytest.mark.gpu


def tol(M, X):
    return 1e-12 * (abs(M) @ np.abs(X)) + 1e-300


def check(Y, M, X, what, factor=1.0, ref=None):
    """Y == M X to the gate; on failure say where (rows / fields / how many entries)."""
    ref = M @ X if ref is None else ref
    bad = ~(np.abs(Y - ref) <= factor * tol(M, X))
    if bad.any():
        r, f = np.nonzero(bad)
        raise AssertionError("%s: %d of %d entries off; rows %s fields %s; worst |dY| %.3e (nan: %d)" % (
            what, bad.sum(), bad.size, np.unique(r)[:12], np.unique(f)[:12], np.nanmax(np.abs(Y - ref)), np.isnan(Y).sum()))


def oracle_csr(Wo):
    return sp.csr_matrix((Wo.val, (Wo.row, Wo.col)), shape=Wo.shape)


@pytest.fixture(scope="module")
def mats(cfg2):
    go, gg = make_oracle(cfg2), make_gpu(cfg2)
    em = cfg2["elevmask"][0]
    rm = gg.regrid_matrices(cfg2["sheets"][0].name, em, scale=True, correctA=True)
    out = {}
    for spec in ("EvI", "AvI"):
        out[spec] = (rm.matrix(spec), oracle_csr(go.matrix(0, em, spec, scale=True, correctA=True)))
    out["_keep"] = (gg, rm)
    return out


@pytest.mark.parametrize("nvar", [64, 66, 200, 512, 700])
def test_stacked_EvI_AvI_vs_oracle(mats, nvar):
    """Full and ragged 64-field strips, one and two field tiles (700 > 512)."""
    from icebin_b200 import api
    (EvI, oE), (AvI, oA) = mats["EvI"], mats["AvI"]
    X = np.random.default_rng(nvar).uniform(-1, 1, (EvI.shape[1], nvar))
    YE, YA = api.apply_many([EvI, AvI], X, layout="cell")
    check(YE, oE, X, "EvI (stacked)")
    check(YA, oA, X, "AvI (stacked)")
    # order of the matrices in the call does not matter; every call gives the same bits
    YA2, YE2 = api.apply_many([AvI, EvI], X, layout="cell")
    assert np.array_equal(YE, YE2) and np.array_equal(YA, YA2)
    # EvI alone takes the same kernel with two running sums
    Y1 = EvI.apply(X, layout="cell")
    check(Y1, oE, X, "EvI alone")
    assert np.array_equal(Y1, EvI.apply(X, layout="cell"))


def test_stacked_conservation(mats):
    """sum wM * (M f) == sum Mw * f for both halves of the stacked apply (test_conserv.py:174-205)."""
    from icebin_b200 import api
    (EvI, _), (AvI, _) = mats["EvI"], mats["AvI"]
    X = np.random.default_rng(5).uniform(0, 1, (EvI.shape[1], 64))
    YE, YA = api.apply_many([EvI, AvI], X, layout="cell")
    for M, Y in ((EvI, YE), (AvI, YA)):
        wM, Mw = M.get_weights(0), M.get_weights(1)
        lhs, rhs = (Y * wM[:, None]).sum(0), (X * Mw[:, None]).sum(0)
        assert np.all(np.abs(lhs - rhs) <= 1e-12 * (np.abs(X) * Mw[:, None]).sum(0))


def test_nan_reaches_only_the_rows_that_name_it(mats):
    """A NaN in one ice cell propagates to the E and A cells that cell feeds, and to no other (structural zeros of the
    stacked records are never multiplied into a live sum)."""
    from icebin_b200 import api
    (EvI, oE), (AvI, oA) = mats["EvI"], mats["AvI"]
    X = np.random.default_rng(9).uniform(-1, 1, (EvI.shape[1], 64))
    bad = [17, 4242, EvI.shape[1] - 1]
    X[bad, 3] = np.nan
    YE, YA = api.apply_many([EvI, AvI], X, layout="cell")
    for Y, o in ((YE, oE), (YA, oA)):
        hit = np.zeros(o.shape[0], bool)
        hit[o.tocsc()[:, bad].tocoo().row] = True
        assert np.array_equal(np.isnan(Y[:, 3]), hit)
        assert not np.isnan(np.delete(Y, 3, axis=1)).any()


def random_wide(rng, nrow1, nrow2, ncol, empty_rows=True):
    """M1 with 0..5 nonzeros per column, M2 with 0..2; some rows and columns empty."""
    def one(nrow, kmax, skip):
        rows, cols = [], []
        live = np.arange(nrow)
        if skip:
            live = live[live % 7 != 3]
        for c in range(ncol):
            k = rng.integers(0, kmax + 1) if c % 11 else 0
            if k:
                # neighbouring rows, as elevation classes are
                r0 = rng.integers(0, len(live))
                rr = live[(r0 + np.arange(k) * rng.integers(1, 3)) % len(live)]
                rows += list(np.unique(rr)); cols += [c] * len(np.unique(rr))
        v = rng.uniform(-1, 1, len(rows))
        return sp.csr_matrix((v, (rows, cols)), shape=(nrow, ncol))
    return one(nrow1, 5, empty_rows), one(nrow2, 2, empty_rows)


@pytest.mark.parametrize("seed,nvar", [(0, 64), (1, 130), (2, 512)])
def test_generic_matrices(seed, nvar):
    """ib_w_create_csr matrices with irregular columns: 0..5 nonzeros, empty rows (zero-filled by the kernel), empty
    columns, record chains that are not elevation-class pairs."""
    from icebin_b200 import api
    rng = np.random.default_rng(seed)
    M1, M2 = random_wide(rng, 300, 40, 20000)
    W1 = api.Weighted.from_csr(M1.shape[0], M1.shape[1], M1.indptr, M1.indices, M1.data)
    W2 = api.Weighted.from_csr(M2.shape[0], M2.shape[1], M2.indptr, M2.indices, M2.data)
    X = rng.uniform(-1, 1, (M1.shape[1], nvar))
    Y1, Y2 = api.apply_many([W1, W2], X, layout="cell")
    check(Y1, M1, X, "M1 (stacked)")
    check(Y2, M2, X, "M2 (stacked)")
    empty1 = np.diff(M1.indptr) == 0
    assert empty1.any() and np.all(Y1[empty1] == 0.0)
    Ys = W1.apply(X, layout="cell")
    check(Ys, M1, X, "M1 alone")
    # the other kernels give the same answer to rounding
    api.tune("use_stream", 0)
    try:
        Z1, Z2 = api.apply_many([W1, W2], X, layout="cell")
    finally:
        api.tune("use_stream", 1)
    check(Z1, M1, X, "M1, other kernels vs one-pass", 2.0, Y1)
    check(Z2, M2, X, "M2, other kernels vs one-pass", 2.0, Y2)


def test_graph_replay_and_back_to_back(mats):
    """Completion counters return to zero inside the kernel and the launches are chained with programmatic stream
    serialization: a captured stacked apply replays, back to back, on changing inputs."""
    import torch
    from icebin_b200 import api
    (EvI, _), (AvI, _) = mats["EvI"], mats["AvI"]
    nvar = 128
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        X = torch.rand((EvI.shape[1], nvar), dtype=torch.float64, device="cuda")
        ref = [torch.empty((M.shape[0], nvar), dtype=torch.float64, device="cuda") for M in (EvI, AvI)]
        out = [torch.zeros_like(r) for r in ref]
        api.apply_many_dev([EvI, AvI], X.data_ptr(), nvar, nvar, [r.data_ptr() for r in ref], [nvar, nvar], True, st.cuda_stream)
        st.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            for _ in range(3):
                api.apply_many_dev([EvI, AvI], X.data_ptr(), nvar, nvar, [o.data_ptr() for o in out], [nvar, nvar], True, st.cuda_stream)
        for rep in range(3):
            for o in out:
                o.zero_()
            g.replay()
            api.apply_many_dev([EvI, AvI], X.data_ptr(), nvar, nvar, [o.data_ptr() for o in out], [nvar, nvar], True, st.cuda_stream)
            st.synchronize()
            assert torch.equal(out[0], ref[0]) and torch.equal(out[1], ref[1]), rep
            X.mul_(0.5)
            for r in ref:
                r.mul_(0.5)


def test_chained_applies_see_the_previous_result(mats, cfg2):
    """I -> E -> I on one stream: the second apply reads what the first wrote (nothing is read from X before
    griddepcontrol.wait, so programmatic launch never lets a kernel see a half-written input)."""
    import torch
    (EvI, oE), _ = mats["EvI"], mats["AvI"]
    gg, rm = mats["_keep"]
    IvE = rm.matrix("IvE")
    nvar = 64
    X = torch.rand((EvI.shape[1], nvar), dtype=torch.float64, device="cuda")
    YE = torch.empty((EvI.shape[0], nvar), dtype=torch.float64, device="cuda")
    YI = torch.empty((IvE.shape[0], nvar), dtype=torch.float64, device="cuda")
    Z = torch.empty_like(YE)
    torch.cuda.synchronize()
    for rep in range(3):
        EvI.apply_dev(X.data_ptr(), nvar, nvar, YE.data_ptr(), nvar)
        IvE.apply_dev(YE.data_ptr(), nvar, nvar, YI.data_ptr(), nvar)
        EvI.apply_dev(YI.data_ptr(), nvar, nvar, Z.data_ptr(), nvar)
    torch.cuda.synchronize()
    import scipy.sparse as sp
    r, c, v = IvE.to_coo_dense()
    oI = sp.csr_matrix((v, (r, c)), shape=IvE.shape)
    Xh = X.cpu().numpy()
    want = oE @ (oI @ (oE @ Xh))
    assert np.allclose(Z.cpu().numpy(), want, rtol=1e-11, atol=1e-13)
