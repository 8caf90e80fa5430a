"""GPU parity of the sigma smoothing operator (K4) and of the conservation correction against the oracle."""
import numpy as np
import pytest

from conftest import make_gpu, make_oracle

pytestmark = pytest.mark.gpu


def csc_order(row, col, val):
    o = np.lexsort((row, col))
    return row[o], col[o], val[o]


@pytest.mark.parametrize("spec", ["IvE", "IvA"])
@pytest.mark.parametrize("sigma", [(100e3, 100e3, 300.0), (75e3, 120e3, np.inf), (60e3, 60e3, 150.0)])
def test_smoothed_matrix_cfg1(cfg1, spec, sigma):
    go, gg = make_oracle(cfg1), make_gpu(cfg1)
    em = cfg1["elevmask"][0]
    Wo = go.matrix(0, em, spec, scale=True, correctA=True, sigma=sigma)
    Wg = gg.regrid_matrices(cfg1["sheets"][0].name, em, scale=True, correctA=True, sigma=sigma).matrix(spec)
    assert not Wg.conservative and Wg.shape == Wo.shape and Wg.nnz == Wo.nnz
    r, c, v = csc_order(*Wg.to_coo_dense())
    assert np.array_equal(r, Wo.row) and np.array_equal(c, Wo.col)
    np.testing.assert_allclose(v, Wo.val, rtol=1e-12, atol=0)
    np.testing.assert_array_equal(Wg.get_weights(0), Wo.wM)
    np.testing.assert_array_equal(Wg.get_weights(1), Wo.Mw)
    # apply with and without the conservation correction
    X = np.random.default_rng(0).uniform(0, 1, (5, Wo.shape[1]))
    for fc in (False, True):
        Yo, Yg = Wo.apply(X, force_conservation=fc), Wg.apply(X, force_conservation=fc)
        np.testing.assert_allclose(Yg, Yo, rtol=1e-12, atol=0)
    Yg = Wg.apply(X, force_conservation=True)
    np.testing.assert_allclose((Yg * Wo.wM).sum(1), (X * Wo.Mw).sum(1), rtol=1e-12)


def test_smoothed_cfg2(cfg2):
    """The configs[2] operator: Greenland 5 km, sigma = (50 km, 50 km, 100 m)."""
    go, gg = make_oracle(cfg2), make_gpu(cfg2)
    em = cfg2["elevmask"][0]
    sigma = (50e3, 50e3, 100.0)
    Wo = go.matrix(0, em, "IvE", scale=True, correctA=True, sigma=sigma)
    Wg = gg.regrid_matrices(cfg2["sheets"][0].name, em, scale=True, correctA=True, sigma=sigma).matrix("IvE")
    r, c, v = csc_order(*Wg.to_coo_dense())
    assert np.array_equal(r, Wo.row) and np.array_equal(c, Wo.col)
    np.testing.assert_allclose(v, Wo.val, rtol=1e-12, atol=0)
    # the configs[2] batch: 512 fields, with and without the conservation correction, at the 1e-12 gate
    X = np.random.default_rng(1).uniform(0, 1, (512, Wo.shape[1]))
    for fc in (False, True):
        np.testing.assert_allclose(Wg.apply(X, force_conservation=fc), Wo.apply(X, force_conservation=fc), rtol=1e-12, atol=0)
    # fields of both signs: |dY| <= 1e-12 * sum |terms| of the row
    import scipy.sparse as sp
    Xs = np.random.default_rng(2).uniform(-1, 1, (64, Wo.shape[1]))
    A = sp.csr_matrix((np.abs(Wo.val), (Wo.row, Wo.col)), shape=Wo.shape)
    dY = np.abs(Wg.apply(Xs, force_conservation=False) - Wo.apply(Xs, force_conservation=False))
    assert np.all(dY <= 1e-12 * (A @ np.abs(Xs).T).T + 1e-300)


def test_smoothing_with_identity_dimI(cfg1):
    """IceCoupler.cpp:461-463: IvE with sigma on an identity dimI holding all nI cells (masked cells -> empty rows)."""
    from icebin_b200 import api
    go, gg = make_oracle(cfg1), make_gpu(cfg1)
    sh, em = cfg1["sheets"][0], cfg1["elevmask"][0]
    sigma = (100e3, 100e3, 300.0)
    dI, dE = api.SparseSet(identity=sh.nI), api.SparseSet()
    Wg = gg.regrid_matrices(sh.name, em).matrix_d("IvE", (dI, dE), scale=True, correctA=True, sigma=sigma)
    Wo = go.matrix(0, em, "IvE", scale=True, correctA=True, sigma=sigma, dims=(np.arange(sh.nI), np.zeros(0, np.int64)))
    r, c, v = csc_order(*Wg.to_coo_dense())
    assert np.array_equal(r, Wo.row) and np.array_equal(c, Wo.col)
    np.testing.assert_allclose(v, Wo.val, rtol=1e-12, atol=0)


def test_smoothed_apply_nan_reaches_only_named_rows(cfg2):
    """The register-blocked group kernel keeps a dense coefficient table with zeros for the rows that do not name a column;
    a NaN in X must still reach exactly the rows whose neighbourhood names that E cell (structural zeros are not multiplied)."""
    import scipy.sparse as sp
    gg = make_gpu(cfg2)
    em = cfg2["elevmask"][0]
    Wg = gg.regrid_matrices(cfg2["sheets"][0].name, em, scale=True, correctA=True, sigma=(50e3, 50e3, 100.0)).matrix("IvE")
    r, c, v = Wg.to_coo_dense()
    M = sp.csr_matrix((v, (r, c)), shape=Wg.shape)
    X = np.random.default_rng(3).uniform(0, 1, (64, Wg.shape[1]))
    bad = [5, 400, Wg.shape[1] - 1]
    X[7, bad] = np.nan
    X[9, bad[0]] = np.inf
    Y = Wg.apply(X, force_conservation=False)
    hit = np.zeros(Wg.shape[0], bool)
    hit[M.tocsc()[:, bad].tocoo().row] = True
    assert np.array_equal(np.isnan(Y[7]), hit)
    hit0 = np.zeros(Wg.shape[0], bool)
    hit0[M.tocsc()[:, [bad[0]]].tocoo().row] = True
    assert np.array_equal(~np.isfinite(Y[9]), hit0)
    ok = np.ones(64, bool); ok[[7, 9]] = False
    Xc = X.copy(); Xc[7, bad] = 0.5; Xc[9, bad[0]] = 0.5
    np.testing.assert_allclose(Y[ok], (M @ Xc[ok].T).T, rtol=1e-12, atol=0)


@pytest.mark.parametrize("nvar", [34, 96, 100, 544])
def test_group_kernel_field_counts_and_row_blocks(cfg2, nvar):
    """The tensor-core group kernel at field counts that leave a consumer warp's 32-field strip partly empty (34, 100),
    fill only some warps (96) or need a second tile (544), with a NaN in the last field; row blocks of the row-block entry
    point agree bit for bit with the whole matrix through the same kernel."""
    import scipy.sparse as sp
    import torch
    gg = make_gpu(cfg2)
    em = cfg2["elevmask"][0]
    Wg = gg.regrid_matrices(cfg2["sheets"][0].name, em, scale=True, correctA=True, sigma=(50e3, 50e3, 100.0)).matrix("IvE")
    r, c, v = Wg.to_coo_dense()
    M = sp.csr_matrix((v, (r, c)), shape=Wg.shape)
    X = np.random.default_rng(nvar).uniform(-1, 1, (nvar, Wg.shape[1]))
    Y = Wg.apply(X, force_conservation=False)
    A = sp.csr_matrix((np.abs(v), (r, c)), shape=Wg.shape)
    assert np.all(np.abs(Y - (M @ X.T).T) <= 1e-12 * (A @ np.abs(X).T).T + 1e-300)
    X[nvar - 1, 17] = np.nan
    Yn = Wg.apply(X, force_conservation=False)
    hit = np.zeros(Wg.shape[0], bool)
    hit[M.tocsc()[:, [17]].tocoo().row] = True
    assert np.array_equal(np.isnan(Yn[nvar - 1]), hit) and np.array_equal(Yn[: nvar - 1], Y[: nvar - 1])
    # row blocks (cell-major device buffers)
    nrow, ncol = Wg.shape
    dX = torch.tensor(np.ascontiguousarray(X.T), device="cuda")
    dYa = torch.full((nrow, nvar), -7.0, dtype=torch.float64, device="cuda")
    dYb = torch.full((nrow, nvar), -7.0, dtype=torch.float64, device="cuda")
    Wg.apply_rows_dev(dX.data_ptr(), nvar, nvar, dYa.data_ptr(), nvar, 0, nrow)
    cut = (nrow // 3) // 32 * 32
    for lo, hi in ((0, cut), (cut, 2 * cut + 64), (2 * cut + 64, nrow)):
        Wg.apply_rows_dev(dX.data_ptr(), nvar, nvar, dYb.data_ptr(), nvar, lo, hi)
    torch.cuda.synchronize()
    assert torch.equal(dYa.isnan(), dYb.isnan()) and torch.equal(dYa.nan_to_num(0.0), dYb.nan_to_num(0.0))
    assert np.array_equal(np.isnan(dYa.cpu().numpy().T), np.isnan(Yn))
