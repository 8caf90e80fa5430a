"""GPU parity of the batched SpMM apply (K3) against the oracle's Eigen-order CSC x dense product."""
import numpy as np
import pytest

from conftest import make_gpu, make_oracle

pytestmark = pytest.mark.gpu


def ref_tol(Wo, X):
    """|dY| <= 1e-12 * sum|terms|: scale for each output row."""
    A = np.zeros(Wo.shape)
    np.add.at(A, (Wo.row, Wo.col), np.abs(Wo.val))
    return 1e-12 * (np.abs(X) @ A.T) + 1e-300


@pytest.mark.parametrize("spec", ["EvI", "AvI", "IvE", "IvA", "EvA", "AvE"])
@pytest.mark.parametrize("nvar", [1, 3, 26, 64])
def test_apply_cfg1(cfg1, spec, nvar):
    go, gg = make_oracle(cfg1), make_gpu(cfg1)
    em = cfg1["elevmask"][0]
    Wo = go.matrix(0, em, spec, scale=True, correctA=True)
    Wg = gg.regrid_matrices(cfg1["sheets"][0].name, em, scale=True, correctA=True).matrix(spec)
    X = np.random.default_rng(nvar).uniform(-1, 1, (nvar, Wo.shape[1]))
    Yo = Wo.apply(X)
    Yf = Wg.apply(X, layout="field")
    Yc = Wg.apply(np.ascontiguousarray(X.T), layout="cell").T
    assert np.array_equal(Yf, Yc)
    assert np.all(np.abs(Yf - Yo) <= ref_tol(Wo, X))


def test_apply_cfg2_bitexact_short_rows(cfg2):
    """Rows with <= 32 nonzeros are accumulated in the oracle's order with mul-then-add: bit-identical.
    (EvI takes the pair-grouped column plan -- every X row read once for its two elevation classes -- whose rows are
    sums of per-chunk partials: deterministic, but associated differently, so it is held to the 1e-12 gate only.)"""
    go, gg = make_oracle(cfg2), make_gpu(cfg2)
    em = cfg2["elevmask"][0]
    for spec in ("IvE", "IvA", "EvI", "AvI"):
        Wo = go.matrix(0, em, spec, scale=True, correctA=True)
        Wg = gg.regrid_matrices(cfg2["sheets"][0].name, em, scale=True, correctA=True).matrix(spec)
        X = np.random.default_rng(7).uniform(-1, 1, (8, Wo.shape[1]))
        Yo, Yg = Wo.apply(X), Wg.apply(X)
        assert np.all(np.abs(Yg - Yo) <= ref_tol(Wo, X)), spec
        rowlen = np.bincount(Wo.row, minlength=Wo.shape[0])
        short = rowlen <= 32
        if spec != "EvI":
            assert np.array_equal(Yg[:, short], Yo[:, short]), spec
        assert np.array_equal(Yg, Wg.apply(X)), spec + ": apply is not deterministic"
        # conservation: sum wM * (M f) == sum Mw * f     (test_conserv.py:174-205)
        lhs, rhs = (Yg * Wo.wM).sum(1), (X * Wo.Mw).sum(1)
        scale = (np.abs(X) * Wo.Mw).sum(1)
        assert np.all(np.abs(lhs - rhs) <= 1e-12 * scale), spec


def seq_rows_reference(M, X):
    """Y = M X with every row summed over ascending columns, one mul and one add at a time (no fma): the order of
    Eigen's column-major sparse * dense (and of the oracle).  M: (rowptr, col, val) CSR, X: [ncol, nvar]."""
    rowptr, col, val = M
    n = np.diff(rowptr)
    Y = np.zeros((len(n), X.shape[1]))
    for k in range(int(n.max()) if len(n) else 0):
        rows = np.nonzero(n > k)[0]
        j = rowptr[rows] + k
        Y[rows] = Y[rows] + val[j, None] * X[col[j]]
    return Y


@pytest.mark.parametrize("spec", ["IvE", "IvA", "EvA"])
@pytest.mark.parametrize("nvar", [32, 64, 100, 200, 300, 512])
def test_apply_distinct_column_groups_bitexact(cfg2, spec, nvar):
    """spmm_gc_kernel (rows of <= 8 nonzeros grouped by their distinct columns, every X row fetched once per group):
    full and partial 64-field slices and a padded leading dimension.  Bit-identical to the sequential row sums over
    the matrix's own values, and (IvE, IvA: values bit-equal to the oracle's) to the oracle's apply."""
    import torch
    go, gg = make_oracle(cfg2), make_gpu(cfg2)
    em = cfg2["elevmask"][0]
    Wg = gg.regrid_matrices(cfg2["sheets"][0].name, em, scale=True, correctA=True).matrix(spec)
    X = np.random.default_rng(nvar).uniform(-1, 1, (nvar, Wg.shape[1]))
    Yref = seq_rows_reference(Wg.to_csr(), np.ascontiguousarray(X.T)).T
    Yg = Wg.apply(X)
    assert np.array_equal(Yg, Yref)
    Wo = go.matrix(0, em, spec, scale=True, correctA=True)
    Yo = Wo.apply(X)
    if spec != "EvA":      # EvA sums > 64 duplicates per entry with a warp: values within 1e-12, not bit-equal
        assert np.array_equal(Yg, Yo)
    assert np.all(np.abs(Yg - Yo) <= ref_tol(Wo, X))
    ld = nvar + 6
    Xd = torch.zeros((Wg.shape[1], ld), dtype=torch.float64, device="cuda")
    Xd[:, :nvar] = torch.from_numpy(np.ascontiguousarray(X.T)).cuda()
    Yd = torch.full((Wg.shape[0], ld), -7.0, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    Wg.apply_dev(Xd.data_ptr(), ld, nvar, Yd.data_ptr(), ld)
    torch.cuda.synchronize()
    Yh = Yd.cpu().numpy()
    assert np.array_equal(Yh[:, :nvar].T, Yref) and np.all(Yh[:, nvar:] == -7.0)


@pytest.mark.parametrize("spec", ["EvI", "AvI"])
@pytest.mark.parametrize("nvar", [64, 200, 512])
def test_apply_wide_matrices_many_fields(cfg2, spec, nvar):
    """EvI (pair-grouped column plan) and AvI (chunked long rows) with full and ragged field tiles against the oracle,
    within 1e-12 of the sum of the magnitudes of a row's terms; deterministic from call to call."""
    import scipy.sparse
    go, gg = make_oracle(cfg2), make_gpu(cfg2)
    em = cfg2["elevmask"][0]
    Wo = go.matrix(0, em, spec, scale=True, correctA=True)
    Wg = gg.regrid_matrices(cfg2["sheets"][0].name, em, scale=True, correctA=True).matrix(spec)
    X = np.random.default_rng(nvar).uniform(-1, 1, (Wo.shape[1], nvar))
    Yo = Wo.apply(np.ascontiguousarray(X.T)).T
    Yg = Wg.apply(X, layout="cell")
    A = scipy.sparse.csr_matrix((np.abs(Wo.val), (Wo.row, Wo.col)), shape=Wo.shape)
    assert np.all(np.abs(Yg - Yo) <= 1e-12 * (A @ np.abs(X)) + 1e-300)
    assert np.array_equal(Yg, Wg.apply(X, layout="cell"))


def test_distinct_column_groups_graph_replay(cfg2):
    """The group kernel hands its units out through a device counter that is never reset (the last warp records where
    the next launch starts), so a captured apply can be replayed, mixed with eager applies and row blocks."""
    import torch
    gg = make_gpu(cfg2)
    M = gg.regrid_matrices(cfg2["sheets"][0].name, cfg2["elevmask"][0], scale=True, correctA=True).matrix("IvE")
    nvar = 192
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        X = torch.rand((M.shape[1], nvar), dtype=torch.float64, device="cuda")
        Y0 = torch.empty((M.shape[0], nvar), dtype=torch.float64, device="cuda")
        Y = torch.zeros_like(Y0)
        M.apply_dev(X.data_ptr(), nvar, nvar, Y0.data_ptr(), nvar, True, st.cuda_stream)        # builds the plan
        st.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            M.apply_dev(X.data_ptr(), nvar, nvar, Y.data_ptr(), nvar, True, st.cuda_stream)
            M.apply_dev(X.data_ptr(), nvar, nvar, Y.data_ptr(), nvar, True, st.cuda_stream)
        for rep in range(3):
            Y.zero_()
            g.replay()
            M.apply_rows_dev(X.data_ptr(), nvar, nvar, Y.data_ptr(), nvar, 64, 4096, st.cuda_stream)
            st.synchronize()
            assert torch.equal(Y, Y0), rep
            X.mul_(0.5); Y0.mul_(0.5)        # exact scaling: the replay must read the new X


def test_apply_M_sparse_and_weights(cfg1):
    go, gg = make_oracle(cfg1), make_gpu(cfg1)
    sh = cfg1["sheets"][0]
    em = cfg1["elevmask"][0]
    Wg = gg.regrid_matrices(sh.name, em).matrix("AvI")
    Wo = go.matrix(0, em, "AvI", scale=True, correctA=True)
    f_s = np.random.default_rng(3).uniform(0, 1, (2, sh.nI))
    B = Wg.apply_M(f_s, fill=-999.0)
    assert B.shape == (2, gg.nA)
    Yo = Wo.apply(f_s[:, Wo.dims[1]])
    assert np.allclose(B[:, Wo.dims[0]], Yo, rtol=1e-12, atol=0)
    untouched = np.ones(gg.nA, bool); untouched[Wo.dims[0]] = False
    assert np.all(B[:, untouched] == -999.0)
    assert np.allclose(Wg.apply_Mw(f_s), (f_s[:, Wo.dims[1]] * Wo.Mw).sum(1), rtol=1e-12)
    assert np.allclose(Wg.apply_wM(B), (B[:, Wo.dims[0]] * Wo.wM).sum(1), rtol=1e-12)
    coo = Wg.to_coo()
    assert coo.shape == (gg.nA, sh.nI) and coo.nnz == Wo.nnz
    one = Wg.apply_M(f_s[0])
    assert one.shape == (gg.nA,) and np.isnan(one[untouched]).all()


def test_constant_roundtrips(cfg2):
    """tests/test_conserv/test_conserv.py:132-172: I->A->I, I->E->I, E->A->E, A->E->A of a constant return it."""
    gg = make_gpu(cfg2)
    rm = gg.regrid_matrices(cfg2["sheets"][0].name, cfg2["elevmask"][0], scale=True, correctA=False)
    M = {s: rm.matrix(s) for s in ("AvI", "IvA", "EvI", "IvE", "EvA", "AvE")}
    for a, b in (("AvI", "IvA"), ("EvI", "IvE"), ("AvE", "EvA"), ("EvA", "AvE")):
        assert np.array_equal(M[a].dim(0), M[b].dim(1)) and np.array_equal(M[a].dim(1), M[b].dim(0))
        one = np.ones((1, M[a].shape[1]))
        back = M[b].apply(M[a].apply(one))
        assert np.abs(back - 1).max() < 1e-13, (a, b)


@pytest.mark.parametrize("spec,nvar", [("IvE", 64), ("AvI", 64), ("EvI", 6), ("IvA", 512)])
def test_row_block_sharding(cfg2, spec, nvar):
    """Output row blocks computed separately (as N GPUs would, inputs replicated, no collective) reproduce the
    full apply bit for bit."""
    import torch
    gg = make_gpu(cfg2)
    M = gg.regrid_matrices(cfg2["sheets"][0].name, cfg2["elevmask"][0], scale=True, correctA=True).matrix(spec)
    dev = torch.device("cuda")
    X = torch.rand((M.shape[1], nvar), dtype=torch.float64, device=dev)
    Yfull = torch.empty((M.shape[0], nvar), dtype=torch.float64, device=dev)
    if spec in ("EvI", "AvI"):
        # the full apply of a wide matrix takes the one-pass kernel (sums associated differently); row blocks share the
        # row-chunk plan with the full range [0, nrow)
        M.apply_rows_dev(X.data_ptr(), nvar, nvar, Yfull.data_ptr(), nvar, 0, M.shape[0])
    else:
        M.apply_dev(X.data_ptr(), nvar, nvar, Yfull.data_ptr(), nvar)
    Y = torch.full((M.shape[0], nvar), float("nan"), dtype=torch.float64, device=dev)
    nparts = 4
    per = -(-M.shape[0] // nparts)
    per = -(-per // 32) * 32
    for k in range(nparts):
        lo, hi = min(k * per, M.shape[0]), min((k + 1) * per, M.shape[0])
        M.apply_rows_dev(X.data_ptr(), nvar, nvar, Y.data_ptr(), nvar, lo, hi)
    torch.cuda.synchronize()
    assert torch.equal(Y, Yfull)
    from icebin_b200 import api
    with pytest.raises(api.IceBinError):
        M.apply_rows_dev(X.data_ptr(), nvar, nvar, Y.data_ptr(), nvar, 5, 64)


def test_apply_many_matches_single_applies(cfg2):
    """ib_w_apply_many: EvI and AvI on the same fields.  The pair is applied in one pass over X (apply_stream.cu), whose
    sums are associated differently from the one-at-a-time kernels': equal to rounding, and deterministic."""
    import scipy.sparse
    from icebin_b200 import api
    gg = make_gpu(cfg2)
    rm = gg.regrid_matrices(cfg2["sheets"][0].name, cfg2["elevmask"][0], scale=True, correctA=True)
    EvI, AvI = rm.matrix("EvI"), rm.matrix("AvI")
    X = np.random.default_rng(0).uniform(-1, 1, (EvI.shape[1], 64))
    absX = np.abs(X)
    for rep in range(3):
        YE, YA = api.apply_many([EvI, AvI], X, layout="cell")
        for M, Y in ((EvI, YE), (AvI, YA)):
            r, c, v = M.to_coo_dense()
            A = scipy.sparse.csr_matrix((np.abs(v), (r, c)), shape=M.shape)
            assert np.all(np.abs(Y - M.apply(X, layout="cell")) <= 2e-12 * (A @ absX) + 1e-300)
        if rep:
            assert np.array_equal(YE, YE0) and np.array_equal(YA, YA0)
        YE0, YA0 = YE, YA
    YEf, YAf = api.apply_many([EvI, AvI], np.ascontiguousarray(X.T), layout="field")
    assert np.array_equal(YEf.T, YE) and np.array_equal(YAf.T, YA)
    # three matrices: forked streams, identical to the single applies
    api.tune("use_stream", 0)
    try:
        YE1, YA1 = api.apply_many([EvI, AvI], X, layout="cell")
        assert np.array_equal(YE1, EvI.apply(X, layout="cell")) and np.array_equal(YA1, AvI.apply(X, layout="cell"))
    finally:
        api.tune("use_stream", 1)
    with pytest.raises(api.IceBinError):
        api.apply_many([EvI, rm.matrix("IvE")], X, layout="cell")      # different input spaces
