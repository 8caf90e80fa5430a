"""GPU parity of the exchange-grid generation (exgrid.cu) against the oracle, bit for bit, and the regrid matrices built
from a generated exchange grid (the generator feeds the path: SURVEY 8f rank 4)."""
import numpy as np
import pytest

from exgrid_common import rect_polys, warped_mesh

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("major_x", [True, False])
def test_exgrid_bitexact_vs_oracle(major_x):
    from icebin_b200 import api
    from oracle import orc
    polys, iA = warped_mesh(ncx=9, ncy=6, nsub=8, seed=3)
    xe = np.linspace(0, 900e3, 301); ye = np.cumsum(np.r_[0.0, np.random.default_rng(1).uniform(2e3, 4e3, 240)])     # ragged rows
    ijo, ao = orc.exgrid_rectilinear(polys, iA, xe, ye, major_x)
    ijg, ag = api.exgrid_rectilinear(polys, iA, xe, ye, major_x)
    assert np.array_equal(ijg, ijo) and np.array_equal(ag, ao)
    # polygons partly or wholly outside the ice grid, and a degenerate sliver
    polys2 = [p + np.array([-500e3, 0.0]) for p in polys[:12]] + [np.array([[10.0, 10.0], [20.0, 10.0], [30.0, 10.0]])]
    iA2 = np.arange(len(polys2))
    ijo, ao = orc.exgrid_rectilinear(polys2, iA2, xe, ye, major_x)
    ijg, ag = api.exgrid_rectilinear(polys2, iA2, xe, ye, major_x)
    assert np.array_equal(ijg, ijo) and np.array_equal(ag, ao) and not np.any(ijg[:, 0] == len(polys2) - 1)
    with pytest.raises(api.IceBinError):
        api.exgrid_rectilinear([np.zeros((2, 2))], [0], xe, ye)


def test_generated_exchange_grid_feeds_the_matrix_build():
    """Rectangular GCM cells: the generated grid equals the synthetic one cell for cell, and AvI built from it conserves."""
    from icebin_b200 import api, synth
    polys, iA, xe, ye = rect_polys("greenland_20km")
    ij, area = api.exgrid_rectilinear(polys, iA, xe, ye, True)
    gA, sh = synth.make_gcm_and_sheet("greenland_20km")
    o1 = np.lexsort((ij[:, 0], ij[:, 1])); o2 = np.lexsort((sh.ex_ij[:, 0], sh.ex_ij[:, 1]))
    assert np.array_equal(ij[o1], sh.ex_ij[o2])
    np.testing.assert_allclose(area[o1], sh.ex_area[o2], rtol=1e-12)
    sh.ex_ij, sh.ex_area = np.ascontiguousarray(ij), area          # A-major order, as the reference's generator emits
    hc = synth.hcdefs_conserv(40)
    em = synth.make_elevmask(sh, hc[-1], 0)
    g = api.GCMRegridder(gA, hc, True)
    g.add_sheet(sh.name, sh)
    M = g.regrid_matrices(sh.name, em, scale=True, correctA=False).matrix("AvI")
    f = np.random.default_rng(0).uniform(0, 1, (3, M.shape[1]))
    lhs, rhs = (M.apply(f) * M.get_weights(0)).sum(1), (f * M.get_weights(1)).sum(1)
    assert np.all(np.abs(lhs / rhs - 1) < 1e-12)
    np.testing.assert_allclose(M.get_weights(1), sh.dx * sh.dx, rtol=1e-12)        # every unmasked ice cell fully tiled
