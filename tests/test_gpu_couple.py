"""GPU parity of the coupling step's regridding side against the oracle: variable transformation in
front of a regrid apply (IceCoupler.cpp:203-252, :404-446) and the E1vE0c correction (e1ve0.cpp:55-106)."""
import numpy as np
import pytest

from conftest import make_gpu, make_oracle
from couple_common import stepped_elevmask, two_sheet_config, var_transform

pytestmark = pytest.mark.gpu


def test_vartrans_bitexact_and_nan_isolation():
    import torch
    from icebin_b200 import api
    from oracle import orc
    M, b = var_transform(7, 11, 3)
    O = np.random.default_rng(0).uniform(-3, 3, (4097, 7))
    O[:, 4] = np.nan                                    # an ALLOW_NAN variable
    Zo = orc.vartrans(O, M, b)
    Zg = api.vartrans(O, M, b)
    assert np.array_equal(Zg, Zo, equal_nan=True)
    uses = M.toarray()[4] != 0
    assert np.isnan(Zg[:, uses]).all() and not np.isnan(Zg[:, ~uses]).any()
    # padded leading dimensions on the device
    dO = torch.full((O.shape[0], 9), 1e300, dtype=torch.float64, device="cuda")
    dO[:, :7] = torch.from_numpy(O).cuda()
    dZ = torch.full((O.shape[0], 13), -5.0, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    api.vartrans_dev(dO.data_ptr(), 9, O.shape[0], M, b, dZ.data_ptr(), 13)
    Zh = dZ.cpu().numpy()
    assert np.array_equal(Zh[:, :11], Zo, equal_nan=True) and np.all(Zh[:, 11:] == -5.0)
    with pytest.raises(api.IceBinError):
        import scipy.sparse
        api.vartrans(O, scipy.sparse.csc_matrix(np.ones((8, 2))), np.zeros(2))       # M names a variable O does not have


def xue_pair(cfg, em_list):
    from icebin_b200 import api
    go, gg = make_oracle(cfg), make_gpu(cfg)
    Xo, Xg = [], []
    for i, sh in enumerate(cfg["sheets"]):
        Xo.append(go.matrix(i, em_list[i], "XvE", scale=False, correctA=True, dims=(np.arange(sh.nX), None)))
        rm = gg.regrid_matrices(sh.name, em_list[i], scale=False, correctA=True)
        Xg.append(rm.matrix_d("XvE", (api.SparseSet(identity=sh.nX), api.SparseSet()), scale=False, correctA=True))
    return Xo, Xg


def test_e1ve0c_two_sheets():
    from icebin_b200 import api
    from oracle import orc
    cfg = two_sheet_config()
    nE = cfg["gridA"].nA * len(cfg["hcdefs"])
    em0 = cfg["elevmask"]
    em1 = [stepped_elevmask(em0[i], 10 + i) for i in range(2)]
    Xo0, Xg0 = xue_pair(cfg, em0)
    Xo1, Xg1 = xue_pair(cfg, em1)
    for a, b in zip(Xo0 + Xo1, Xg0 + Xg1):          # the inputs themselves agree (XvE unscaled, identity X dim)
        assert a.shape == b.shape and a.nnz == b.nnz and np.array_equal(a.dims[1], b.dim(1))
    ro, co, vo = orc.e1ve0c(Xo1, Xo0, nE)
    rg, cg, vg = api.e1ve0c(Xg1, Xg0, nE)
    assert np.array_equal(rg, ro) and np.array_equal(cg, co)          # pattern and order, bit-exact
    assert np.abs(vg - vo).max() <= 1e-12 * np.abs(vo).max()
    # unchanged elevation classes => the correction vanishes identically
    rz, cz, vz = api.e1ve0c(Xg0, Xg0, nE)
    assert len(vz) > 0 and np.all(vz == 0.0)
    with pytest.raises(api.IceBinError):
        api.e1ve0c(Xg1, Xg0[::-1], nE)                                  # sheets with different exchange grids


def test_couple_step_device_pipeline(cfg2):
    """ice -> gcm: gcm_ivalsX = X1vI * (ice_ovalsI * gvi.M + gvi.b) for X in {A (unscaled, correctA), E (unscaled, no
    correctA)} applied together, then gcm -> ice with the smoothed-free IvE: ice_ivalsI = IvE0 * (gcm_ovalsE0 * ivg.M + b);
    everything stays on the device between the transformation and the regrid (IceCoupler.cpp:361-468, :203-252)."""
    import torch
    from icebin_b200 import api
    from oracle import orc
    go, gg = make_oracle(cfg2), make_gpu(cfg2)
    sh, em = cfg2["sheets"][0], cfg2["elevmask"][0]
    nI = sh.nI
    rm = gg.regrid_matrices(sh.name, em)
    dimI = api.SparseSet(identity=nI)
    dimE, dimA = api.SparseSet(), api.SparseSet()
    E1vI = rm.matrix_d("EvI", (dimE, dimI), scale=False, correctA=False)
    A1vI = rm.matrix_d("AvI", (dimA, dimI), scale=False, correctA=True)
    oE = go.matrix(0, em, "EvI", scale=False, correctA=False, dims=(None, np.arange(nI)))
    oA = go.matrix(0, em, "AvI", scale=False, correctA=True, dims=(None, np.arange(nI)))
    m, n = 5, 8
    Mt, b = var_transform(m, n, 2)
    O = np.random.default_rng(4).uniform(0, 10, (nI, m))
    Zo = orc.vartrans(O, Mt, b)
    dO = torch.from_numpy(O).cuda()
    dZ = torch.empty((nI, n), dtype=torch.float64, device="cuda")
    dYE = torch.empty((E1vI.shape[0], n), dtype=torch.float64, device="cuda")
    dYA = torch.empty((A1vI.shape[0], n), dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    api.vartrans_dev(dO.data_ptr(), m, nI, Mt, b, dZ.data_ptr(), n)
    api.apply_many_dev([E1vI, A1vI], dZ.data_ptr(), n, n, [dYE.data_ptr(), dYA.data_ptr()], [n, n])
    torch.cuda.synchronize()
    for Yg, Wo in ((dYE.cpu().numpy(), oE), (dYA.cpu().numpy(), oA)):
        Yo = Wo.apply(np.ascontiguousarray(Zo.T)).T
        import scipy.sparse
        A = scipy.sparse.csr_matrix((np.abs(Wo.val), (Wo.row, Wo.col)), shape=Wo.shape)
        assert np.all(np.abs(Yg - Yo) <= 1e-12 * (A @ np.abs(Zo)) + 1e-300)
    # gcm -> ice
    IvE = rm.matrix_d("IvE", (dimI, dimE), scale=True, correctA=True)
    oIvE = go.matrix(0, em, "IvE", scale=True, correctA=True, dims=(np.arange(nI), oE.dims[0]))
    G = np.random.default_rng(5).uniform(-1, 1, (IvE.shape[1], n))
    Mt2, b2 = var_transform(n, 4, 7)
    Zo2 = orc.vartrans(G, Mt2, b2)
    Yo2 = oIvE.apply(np.ascontiguousarray(Zo2.T)).T
    dG = torch.from_numpy(G).cuda()
    dZ2 = torch.empty((IvE.shape[1], 4), dtype=torch.float64, device="cuda")
    dY2 = torch.empty((nI, 4), dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    api.vartrans_dev(dG.data_ptr(), n, IvE.shape[1], Mt2, b2, dZ2.data_ptr(), 4)
    IvE.apply_dev(dZ2.data_ptr(), 4, 4, dY2.data_ptr(), 4)
    torch.cuda.synchronize()
    assert np.array_equal(dY2.cpu().numpy(), Yo2)          # short rows, mul-then-add: bit-identical
