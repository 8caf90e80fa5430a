"""CPU: the synthetic grid generator produces a consistent exchange grid."""
import numpy as np

from icebin_b200 import synth


def test_exchange_grid_partitions_ice_cells(cfg1):
    sh, gA = cfg1["sheets"][0], cfg1["gridA"]
    assert sh.ex_ij.dtype == np.int32 and sh.ex_ij.shape == (sh.nX, 2)
    areaI = np.bincount(sh.ex_ij[:, 1], sh.ex_area, sh.nI)
    np.testing.assert_allclose(areaI, sh.dx * sh.dx, rtol=1e-12)          # every ice cell is exactly tiled
    assert np.all(sh.ex_area > 0)
    assert np.all(np.diff(sh.ex_ij[:, 1]) >= 0)                            # I-major emission order
    assert set(np.unique(sh.ex_ij[:, 0])) <= set(gA.dim)                   # only realised A cells are referenced
    areaA = np.bincount(sh.ex_ij[:, 0], sh.ex_area, gA.nA)
    assert areaA.max() <= synth.A_W * synth.A_H * (1 + 1e-12)
    splits = np.bincount(np.bincount(sh.ex_ij[:, 1]))
    assert splits[1] > 0 and splits[2] > 0 and splits[4] > 0              # 1/2/4-way splits all occur


def test_configs_shapes():
    c2 = synth.config(2)
    sh = c2["sheets"][0]
    assert (sh.nx, sh.ny, sh.nI) == (301, 561, 168861) and len(c2["hcdefs"]) == 40
    frac = np.mean(~np.isnan(c2["elevmask"][0]))
    assert 0.38 < frac < 0.43
    assert np.nanmax(c2["elevmask"][0]) <= c2["hcdefs"][-1]
