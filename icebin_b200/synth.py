"""Synthetic grids of the shapes BASELINE.json names.

The reference builds its exchange grid with CGAL polygon overlap (gridgen/GridGen_Exchange.cpp:210-284),
which is upstream of the hot path and out of scope; here GCM cells are modelled as axis-aligned rectangles in
the ice-grid plane so the A-I overlap is a product of 1-D interval overlaps.  Everything is seeded and
vectorised (numpy) so the 36 M-cell Antarctic grid is generated in seconds.

Index conventions follow the reference:
  * A grid: ModelE lon-lat, i fastest: iA = j*im + i              (GridSpec_LonLat indices {1,0}, GridSpec.cpp:117)
  * I grid: PISM order, x slowest:     iI = ix*ny + iy            (gridgen/searise_grid.cpp:104-108)
  * ExchangeGrid: interleaved (iA, iI) int32 pairs + overlap area (AbbrGrid.hpp:40-78), emitted I-major
"""
from dataclasses import dataclass, field

import numpy as np

KM = 1000.0


@dataclass
class GridA:
    """What the hot path reads of the GCM AbbrGrid (AbbrGrid.hpp:93-109)."""
    im: int
    jm: int
    nA: int                       # sparse extent
    dim: np.ndarray               # int64 [n_real]: realised cells, dense -> sparse
    native_area: np.ndarray       # float64 [n_real]


@dataclass
class SheetGrid:
    """What the hot path reads of one IceRegridder (IceRegridder.hpp:60-133)."""
    name: str
    nx: int
    ny: int
    dx: float
    nI: int
    dimI: np.ndarray              # int64 [nI] (identity: every ice cell is realised)
    native_areaI: np.ndarray      # float64 [nI]
    centroid_xy: np.ndarray       # float64 [nI, 2]
    ex_ij: np.ndarray             # int32 [nX, 2]  (iA, iI)
    ex_area: np.ndarray           # float64 [nX]
    gridA_proj_area: np.ndarray   # float64 [n_real_A]  (IceRegridder.cpp:93-119; an input here)
    interp_style: int = 0         # 0 = Z_INTERP, 1 = ELEV_CLASS_INTERP
    extra: dict = field(default_factory=dict)

    @property
    def nX(self):
        return len(self.ex_area)


# SeaRISE-shaped grids (gridgen/searise_grid.cpp:111-123): (nx, ny, dx_km)
SHEET_SHAPES = {
    "greenland_50km": (31, 57, 50.0),
    "greenland_20km": (76, 141, 20.0),
    "greenland_5km": (301, 561, 5.0),
    "greenland_1km": (1501, 2801, 1.0),
    "antarctica_5km": (1201, 1201, 5.0),
    "antarctica_1km": (6000, 6000, 1.0),
}

A_W = 95.0 * KM      # 2.5 degrees of longitude at ~70N
A_H = 222.0 * KM     # 2 degrees of latitude
A_OFF = (-13.7 * KM, -41.3 * KM)   # non-commensurate offset => 1/2/4-way splits of ice cells


def hcdefs_example1(nhc=10):
    """examples/example1/runme.py:77"""
    return np.arange(nhc) * (4000.0 / nhc) - 4000.0 / (2 * nhc)


def hcdefs_conserv(nhc=40):
    """tests/test_conserv/write_icebin_in_base.py:44"""
    return np.arange(nhc) * 100.0 - 50.0


def _pieces_1d(n, d, a, off):
    """1-D overlap of n cells of width d (starting at 0) with cells of width a starting at `off` (<= 0).
    Returns (idx[n,2], len[n,2]) -- each ice cell meets one or two coarse cells."""
    lo = np.arange(n) * d
    hi = lo + d
    p0 = np.floor((lo - off) / a).astype(np.int64)
    edge = off + (p0 + 1) * a                    # upper edge of coarse cell p0
    l0 = np.minimum(hi, edge) - lo
    l1 = np.maximum(hi - edge, 0.0)
    idx = np.stack([p0, p0 + 1], axis=1)
    ln = np.stack([l0, l1], axis=1)
    return idx, ln


def make_gcm_and_sheet(shape="greenland_5km", anchor=(40, 60), im=144, jm=90, name=None, margin=1,
                       gridA=None, cell=None):
    """Build (GridA, SheetGrid).  `anchor` = (i, j) of the A cell holding the ice grid's lower-left corner region.
    Pass an existing `gridA` to add a further sheet to the same GCM grid (its realised set must cover the sheet).
    cell = (width, height) of a GCM cell in the ice-grid plane (default: the 2.5 x 2 degree cell; the ModelE ocean grid of
    GCMRegridder_ModelE is the same grid refined 2:1, cell = (A_W / 2, A_H / 2))."""
    nx, ny, dxkm = SHEET_SHAPES[shape] if isinstance(shape, str) else shape
    dx = dxkm * KM
    A_W, A_H = cell if cell is not None else (globals()["A_W"], globals()["A_H"])
    px, lx = _pieces_1d(nx, dx, A_W, A_OFF[0])
    py, ly = _pieces_1d(ny, dx, A_H, A_OFF[1])
    np_, nq = int(px.max()) + 1, int(py.max()) + 1
    i0, j0 = anchor
    assert i0 + np_ <= im and j0 + nq <= jm, "ice grid does not fit in the GCM grid at this anchor"

    # [nx, ny, sy, sx] in C order == I-major emission, then (q, p) inside one ice cell
    area = lx[:, None, None, :] * ly[None, :, :, None]
    iA = (j0 + py[None, :, :, None]) * im + (i0 + px[:, None, None, :])
    iI = (np.arange(nx)[:, None, None, None] * ny + np.arange(ny)[None, :, None, None])
    keep = area > 0
    ex_area = np.ascontiguousarray(area[keep])
    iA_f = np.broadcast_to(iA, area.shape)[keep]
    iI_f = np.broadcast_to(iI, area.shape)[keep]
    ex_ij = np.empty((len(ex_area), 2), np.int32)
    ex_ij[:, 0] = iA_f
    ex_ij[:, 1] = iI_f

    if gridA is None:
        pp, qq = np.meshgrid(np.arange(-margin, np_ + margin), np.arange(-margin, nq + margin), indexing="xy")
        ii = i0 + pp.ravel()
        jj = j0 + qq.ravel()
        ok = (ii >= 0) & (ii < im) & (jj >= 0) & (jj < jm)
        dimA = np.sort((jj[ok] * im + ii[ok]).astype(np.int64))
        pi = dimA % im - i0
        qi = dimA // im - j0
        ratio = 1.0 + 0.08 * np.sin(0.37 * pi + 0.21 * qi)       # smooth native/proj ratio in [0.92, 1.08]
        gridA = GridA(im, jm, im * jm, dimA, (A_W * A_H) * ratio)
    proj = np.full(len(gridA.dim), A_W * A_H)

    cen = np.stack([np.repeat((np.arange(nx) + 0.5) * dx, ny), np.tile((np.arange(ny) + 0.5) * dx, nx)], axis=1)
    sheet = SheetGrid(name or (shape if isinstance(shape, str) else "sheet"), nx, ny, dx, nx * ny,
                      np.arange(nx * ny, dtype=np.int64), np.full(nx * ny, dx * dx), cen, ex_ij, ex_area, proj)
    return gridA, sheet


def make_elevmask(sheet, top, seed=0, frac=0.72, dome=3200.0, noise=30.0):
    """Elliptical dome: elev = dome*(1-r^2) + N(0, noise) clipped to [0, top]; NaN outside (ElevMask.cpp:56-71)."""
    rng = np.random.default_rng(seed)
    hx, hy = 0.5 * sheet.nx * sheet.dx, 0.5 * sheet.ny * sheet.dx
    x = sheet.centroid_xy[:, 0] - hx
    y = sheet.centroid_xy[:, 1] - hy
    r2 = (x / (frac * hx)) ** 2 + (y / (frac * hy)) ** 2
    elev = dome * (1.0 - r2) + rng.normal(0.0, noise, size=sheet.nI)
    elev = np.clip(elev, 0.0, top)
    elev[r2 >= 1.0] = np.nan
    return elev


def config(n, seed=0):
    """The five BASELINE.json configs -> dict(gridA, sheets[], hcdefs, elevmask[])."""
    if n == 1:
        hc = hcdefs_example1(10)
        gA, sh = make_gcm_and_sheet("greenland_50km")
        return dict(gridA=gA, sheets=[sh], hcdefs=hc, elevmask=[make_elevmask(sh, hc[-1], seed)])
    if n in (2, 3):
        hc = hcdefs_conserv(40)
        gA, sh = make_gcm_and_sheet("greenland_5km")
        return dict(gridA=gA, sheets=[sh], hcdefs=hc, elevmask=[make_elevmask(sh, hc[-1], seed)])
    if n == 5:
        hc = hcdefs_conserv(40)
        # one GCM grid realising the whole globe so both sheets fit
        im, jm = 144, 90
        ids = np.arange(im * jm, dtype=np.int64)
        full = GridA(im, jm, im * jm, ids, (A_W * A_H) * (1.0 + 0.08 * np.sin(0.37 * (ids % im) + 0.21 * (ids // im))))
        _, gr = make_gcm_and_sheet("greenland_1km", anchor=(40, 60), gridA=full, name="greenland")
        _, an = make_gcm_and_sheet("antarctica_1km", anchor=(40, 2), gridA=full, name="antarctica")
        return dict(gridA=full, sheets=[gr, an], hcdefs=hc,
                    elevmask=[make_elevmask(gr, hc[-1], seed), make_elevmask(an, hc[-1], seed + 1)])
    raise ValueError(n)
