"""CPU: the oracle's exchange-grid generation (polygon x rectangle clipping, gridgen/GridGen_Exchange.cpp:175-316) against
the synthetic generator's closed form for rectangles and against the partition identities any exchange grid obeys."""
import numpy as np

from exgrid_common import rect_polys, shoelace, warped_mesh
from icebin_b200 import synth
from oracle import orc


def test_rectangles_reproduce_the_synthetic_exchange_grid():
    polys, iA, xe, ye = rect_polys("greenland_50km")
    ij, area = orc.exgrid_rectilinear(polys, iA, xe, ye, True)
    gA, sh = synth.make_gcm_and_sheet("greenland_50km")
    key = lambda a: np.lexsort((a[:, 0], a[:, 1]))          # by (iI, iA)
    o1, o2 = key(ij), key(sh.ex_ij)
    assert len(area) == sh.nX and np.array_equal(ij[o1], sh.ex_ij[o2])
    np.testing.assert_allclose(area[o1], sh.ex_area[o2], rtol=1e-12)


def test_partition_identities_on_warped_cells():
    polys, iA = warped_mesh()
    nx, ny = 173, 131
    xe = np.linspace(0, 900e3, nx + 1); ye = np.linspace(0, 700e3, ny + 1)
    ij, area = orc.exgrid_rectilinear(polys, iA, xe, ye, True)
    assert np.all(area > 0) and np.all(np.diff(np.nonzero(np.diff(ij[:, 0]))[0]) > 0)
    # (1) every polygon lies inside the ice grid: its exchange cells add up to its area
    for p, a in zip(polys, iA):
        np.testing.assert_allclose(area[ij[:, 0] == a].sum(), shoelace(p), rtol=1e-12)
    # (2) ice cells strictly inside the mapped region are tiled by the polygons
    cover = np.bincount(ij[:, 1], weights=area, minlength=nx * ny).reshape(nx, ny)
    cell = np.outer(np.diff(xe), np.diff(ye))
    inner = cover[nx // 3: 2 * nx // 3, ny // 3: 2 * ny // 3] / cell[nx // 3: 2 * nx // 3, ny // 3: 2 * ny // 3]
    np.testing.assert_allclose(inner, 1.0, rtol=1e-10)
    # (3) ice cells ascending inside a polygon, polygons in input order; the other index convention is a relabelling
    for a in iA:
        assert np.all(np.diff(ij[ij[:, 0] == a, 1]) > 0)
    ij2, area2 = orc.exgrid_rectilinear(polys, iA, xe, ye, False)
    ix, iy = ij[:, 1] // ny, ij[:, 1] % ny
    k1 = np.lexsort((iy * nx + ix, ij[:, 0])); k2 = np.lexsort((ij2[:, 1], ij2[:, 0]))
    assert np.array_equal((iy * nx + ix)[k1], ij2[k2, 1]) and np.array_equal(area[k1], area2[k2])
