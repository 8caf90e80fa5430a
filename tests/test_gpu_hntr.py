"""GPU parity of the Hntr regrid (K5) against the oracle's 4-nested-loop restatement + closed-form identities."""
import numpy as np
import pytest

from icebin_b200 import api
from oracle import orc

pytestmark = pytest.mark.gpu

G = api.GRIDS
PAIRS = [  # (B, A)
    (api.HntrSpec(4, 2, 0.0, 90.0 * 60), api.HntrSpec(8, 4, 0.0, 45.0 * 60)),
    (api.HntrSpec(8, 4, 0.0, 45.0 * 60), api.HntrSpec(16, 8, 0.0, 22.5 * 60)),
    (api.HntrSpec(16, 8, 0.0, 22.5 * 60), api.HntrSpec(8, 4, 0.0, 45.0 * 60)),      # refining
    (G["g2hx2"], G["g1qx1"]), (G["g1qx1"], G["g2hx2"]), (G["g1qx1"], G["ghxh"]), (G["g1qx1"], G["g1x1"]),
    (G["g2hx2"], G["g10mx10m"]), (G["g5x4"], G["g2hx2"]),
    (api.HntrSpec(7, 5, 0.5, 36.0 * 60), api.HntrSpec(9, 11, 0.25, 180.0 * 60 / 11)),   # odd, offset grids
]


def spec_t(s):
    return (s.im, s.jm, s.offi, s.dlat)


@pytest.mark.parametrize("B,A", PAIRS)
def test_partition_tables_bit_equal(B, A):
    hg, ho = api.Hntr(17.17, B, A, 0.0), orc.Hntr(spec_t(B), spec_t(A), 0.0)
    for name in ("SINA", "SINB", "FMIN", "FMAX", "GMIN", "GMAX", "IMIN", "IMAX", "JMIN", "JMAX", "dxypA", "dxypB"):
        assert np.array_equal(hg.get(name), ho.get(name)), name


@pytest.mark.parametrize("B,A", PAIRS)
@pytest.mark.parametrize("mode", ["f64_w", "f64_const", "i16_const", "f64_mask"])
def test_regrid_vs_oracle(B, A, mode):
    rng = np.random.default_rng(A.im * 7 + B.im)
    datmis = -1e30
    hg, ho = api.Hntr(17.17, B, A, datmis), orc.Hntr(spec_t(B), spec_t(A), datmis)
    if mode == "i16_const":
        src = rng.integers(-500, 4000, A.size).astype(np.int16); W = None
    else:
        src = rng.uniform(0, 1, A.size)
        W = {"f64_w": rng.uniform(0.1, 1, A.size), "f64_const": None, "f64_mask": (rng.uniform(0, 1, A.size) < 0.3).astype(float)}[mode]
        if mode == "f64_mask" and A.size > 64:
            W.reshape(A.jm, A.im)[: A.jm // 2] = 0.0     # whole B cells without weight -> DATMIS
    for mp in (False, True):
        Bo = ho.regrid(W, src, mean_polar=mp)
        Bg = hg.regrid(W, src, mean_polar=mp)
        miss = Bo == datmis
        assert np.array_equal(miss, Bg == datmis)
        scale = np.abs(src.astype(float)).max()
        assert np.all(np.abs(Bg[~miss] - Bo[~miss]) <= 1e-12 * scale)


def test_conservation_and_kat():
    """tests/test_hntr.cpp:464-494 (sum4 == sum8) and the i+j field of regrid1 (:430-445) in closed form."""
    g4, g8 = api.HntrSpec(8, 4, 0.0, 45.0 * 60), api.HntrSpec(16, 8, 0.0, 22.5 * 60)
    h = api.Hntr(17.17, g4, g8)
    i, j = np.meshgrid(np.arange(1, 17), np.arange(1, 9), indexing="xy")
    v8 = (i + j).astype(float)
    v4 = h.regrid(None, v8.ravel()).reshape(4, 8)
    d8, d4 = h.get("dxypA"), h.get("dxypB")
    assert abs((v8 * d8[:, None]).sum() / (v4 * d4[:, None]).sum() - 1) < 1e-12
    # exact 2:1 nesting: B = area-weighted mean of the 2x2 block
    blk = v8.reshape(4, 2, 8, 2)
    w = d8.reshape(4, 2)[:, :, None, None]
    expect = (blk * w).sum(axis=(1, 3)) / (2 * w.sum(axis=(1, 3)))
    assert np.allclose(v4, expect, rtol=1e-13)


@pytest.mark.parametrize("B,A", PAIRS[:5])
def test_overlap_matrix(B, A):
    """tests/test_hntr.cpp:332-427: row/col sums of overlap == R^2 dxyp, total 4 pi R^2; entries match the oracle."""
    hg, ho = api.Hntr(17.17, B, A), orc.Hntr(spec_t(B), spec_t(A))
    for R in (1.0, 2.0):
        r, c, v = hg.overlap(R)
        ro, co, vo = ho.matrix(R)
        assert np.array_equal(r, ro) and np.array_equal(c, co) and np.array_equal(v, vo)
        areaB = np.bincount(r, v, B.size); areaA = np.bincount(c, v, A.size)
        assert np.allclose(areaB, np.repeat(hg.get("dxypB"), B.im) * R * R, rtol=1e-12)
        assert np.allclose(areaA, np.repeat(hg.get("dxypA"), A.im) * R * R, rtol=1e-12)
        assert abs(v.sum() / (4 * np.pi * R * R) - 1) < 1e-12
    r, c, v = hg.scaled_regrid_matrix()
    ro, co, vo = ho.matrix(None)
    assert np.array_equal(r, ro) and np.array_equal(c, co) and np.array_equal(v, vo)


def test_errors():
    with pytest.raises(api.IceBinError):
        api.Hntr(17.17, api.HntrSpec(144, 90, 0., 2.0), api.GRIDS["g1qx1"])     # dlat given in degrees
    h = api.Hntr(17.17, api.GRIDS["g2hx2"], api.GRIDS["g1qx1"])
    with pytest.raises(api.IceBinError):
        h.regrid(None, np.zeros(10))
