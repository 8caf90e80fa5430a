"""GPU, BASELINE.json full sizes: size-independent properties (the oracle would take minutes here).
  * Greenland 1 km (1501x2801, nX 4.26 M), 40 ECs: unscaled == diag(wM) scaled, mass conservation, constant round trips,
    dense maps consistent across matrices, weights equal the geometric areas (test_conserv.py:89-205).
  * Hntr at ETOPO1 scale (21600x10800 -> 144x90 and 360x180): conservation of the area integral (test_hntr.cpp:308, 492),
    int16 and fp64 sources agree, constant field stays constant."""
import numpy as np
import pytest
import torch

from icebin_b200 import api, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gr1km():
    hc = synth.hcdefs_conserv(40)
    gA, sh = synth.make_gcm_and_sheet("greenland_1km", anchor=(40, 60))
    em = synth.make_elevmask(sh, hc[-1], 0)
    g = api.GCMRegridder(gA, hc, True)
    g.add_sheet(sh.name, sh)
    return g, gA, sh, em


def test_greenland_1km_properties(gr1km):
    g, gA, sh, em = gr1km
    assert sh.nX == 4264508
    S = {s: g.regrid_matrices(sh.name, em, scale=True, correctA=True).matrix(s) for s in ("AvI", "EvI", "IvE", "IvA", "EvA", "AvE")}
    U = {s: g.regrid_matrices(sh.name, em, scale=False, correctA=True).matrix(s) for s in ("AvI", "EvI")}
    nunmasked = int(np.sum(~np.isnan(em[sh.ex_ij[:, 1]])))
    assert S["AvI"].nnz == nunmasked and S["EvI"].nnz <= 2 * nunmasked and S["EvA"].nnz == S["EvA"].shape[0]
    for a, b in (("AvI", "IvA"), ("EvI", "IvE"), ("EvA", "AvE")):
        assert np.array_equal(S[a].dim(0), S[b].dim(1)) and np.array_equal(S[a].dim(1), S[b].dim(0))
    for s in ("AvI", "EvI"):
        rs, cs, vs = S[s].to_coo_dense()
        ru, cu, vu = U[s].to_coo_dense()
        assert np.array_equal(rs, ru) and np.array_equal(cs, cu)
        np.testing.assert_allclose(vu, S[s].get_weights(0)[rs] * vs, rtol=1e-13)
    # weights are areas: Mw of AvI = covered area of each ice cell = dx^2 (every unmasked cell fully tiled)
    np.testing.assert_allclose(S["AvI"].get_weights(1), sh.dx * sh.dx, rtol=1e-12)
    # conservation and round trips with 8 fields resident on the device
    nvar, dev = 8, torch.device("cuda")
    nI = S["EvI"].shape[1]
    X = torch.rand((nI, nvar), dtype=torch.float64, device=dev)
    X[:, 0] = 1.0
    for s in ("AvI", "EvI"):
        M = S[s]
        Y = torch.empty((M.shape[0], nvar), dtype=torch.float64, device=dev)
        M.apply_dev(X.data_ptr(), nvar, nvar, Y.data_ptr(), nvar)
        torch.cuda.synchronize()
        wM = torch.from_numpy(M.get_weights(0)).to(dev); Mw = torch.from_numpy(M.get_weights(1)).to(dev)
        lhs, rhs = (Y * wM[:, None]).sum(0), (X * Mw[:, None]).sum(0)
        assert torch.all(torch.abs(lhs / rhs - 1) < 1e-12), s
    # correctA puts the native/proj ratio on the A/E side only, so E->I of (I->E of 1) is the ratio-weighted 1;
    # the plain round trip is checked without correctA
    P = {s: g.regrid_matrices(sh.name, em, scale=True, correctA=False).matrix(s) for s in ("EvI", "IvE")}
    one = torch.ones((nI, 1), dtype=torch.float64, device=dev)
    yE = torch.empty((P["EvI"].shape[0], 1), dtype=torch.float64, device=dev)
    zI = torch.empty((nI, 1), dtype=torch.float64, device=dev)
    P["EvI"].apply_dev(one.data_ptr(), 1, 1, yE.data_ptr(), 1)
    P["IvE"].apply_dev(yE.data_ptr(), 1, 1, zI.data_ptr(), 1)
    torch.cuda.synchronize()
    assert float(torch.abs(zI - 1).max()) < 1e-12


def test_greenland_1km_iside_group_kernel(gr1km):
    """IvE / IvA at 1 km (1.74 M rows of <= 8 nonzeros) through spmm_gc_kernel with 64 and 200 fields: conservation
    sum wM (M f) == sum Mw f, linearity M(2f + g) == 2 Mf + Mg to rounding, and agreement of the row-block shards with the
    full apply, bit for bit."""
    g, gA, sh, em = gr1km
    rm = g.regrid_matrices(sh.name, em, scale=True, correctA=True)
    dev = torch.device("cuda")
    for spec in ("IvE", "IvA"):
        M = rm.matrix(spec)
        assert int(np.diff(M.to_csr()[0]).max()) <= 8
        wM = torch.from_numpy(M.get_weights(0)).to(dev); Mw = torch.from_numpy(M.get_weights(1)).to(dev)
        for nvar in (64, 200):
            X = torch.rand((M.shape[1], nvar), dtype=torch.float64, device=dev) + 0.5
            G = torch.rand((M.shape[1], nvar), dtype=torch.float64, device=dev)
            Y = torch.empty((M.shape[0], nvar), dtype=torch.float64, device=dev)
            Yg = torch.empty_like(Y); Yc = torch.empty_like(Y)
            torch.cuda.synchronize()
            M.apply_dev(X.data_ptr(), nvar, nvar, Y.data_ptr(), nvar)
            M.apply_dev(G.data_ptr(), nvar, nvar, Yg.data_ptr(), nvar)
            C = 2 * X + G
            torch.cuda.synchronize()
            M.apply_dev(C.data_ptr(), nvar, nvar, Yc.data_ptr(), nvar)
            torch.cuda.synchronize()
            lhs, rhs = (Y * wM[:, None]).sum(0), (X * Mw[:, None]).sum(0)
            assert torch.all(torch.abs(lhs / rhs - 1) < 1e-12), (spec, nvar)
            assert float((Yc - (2 * Y + Yg)).abs().max()) < 1e-12 * float(Yc.abs().max()), (spec, nvar)
            Yb = torch.full_like(Y, float("nan"))
            cut = (M.shape[0] // 3) // 32 * 32
            for lo, hi in ((0, cut), (cut, 2 * cut), (2 * cut, M.shape[0])):
                M.apply_rows_dev(X.data_ptr(), nvar, nvar, Yb.data_ptr(), nvar, lo, hi)
            torch.cuda.synchronize()
            assert torch.equal(Yb, Y), (spec, nvar)


def test_greenland_1km_deterministic(gr1km):
    g, gA, sh, em = gr1km
    a = g.regrid_matrices(sh.name, em).matrix("EvI")
    b = g.regrid_matrices(sh.name, em).matrix("EvI")
    for x, y in zip(a.to_csr(), b.to_csr()):
        assert np.array_equal(x, y)
    assert np.array_equal(a.dim(0), b.dim(0)) and np.array_equal(a.get_weights(0), b.get_weights(0))


def test_hntr_etopo1_scale():
    A = api.GRIDS["g1mx1m"]
    rng = np.random.default_rng(0)
    src16 = rng.integers(-500, 4000, A.size, dtype=np.int16)
    for bname in ("g2hx2", "g1x1"):
        B = api.GRIDS[bname]
        h = api.Hntr(17.17, B, A)
        out = h.regrid(None, src16)
        dA, dB = h.get("dxypA"), h.get("dxypB")
        sa = (src16.reshape(A.jm, A.im).astype(np.float64).sum(1) * dA).sum()
        sb = (out.reshape(B.jm, B.im).sum(1) * dB).sum()
        assert abs(sa / sb - 1) < 1e-12, bname
        if bname == "g2hx2":
            out64 = h.regrid(None, src16.astype(np.float64))
            np.testing.assert_allclose(out64, out, rtol=1e-13)
            const = h.regrid(None, np.full(A.size, 7, np.int16), mean_polar=True)
            assert np.abs(const - 7).max() < 1e-12
