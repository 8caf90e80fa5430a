"""CPU: pins the Hntr oracle against closed forms (tests/test_hntr.cpp:332-427, 464-494), against an independent
loop-for-loop numpy twin, and against the committed vectors in tests/golden/hntr_kat.npz."""
import os

import numpy as np
import pytest

from oracle import orc, oracle_np

GOLD = os.path.join(os.path.dirname(__file__), "golden")
g2, g4, g8 = (4, 2, 0.0, 90.0 * 60), (8, 4, 0.0, 45.0 * 60), (16, 8, 0.0, 22.5 * 60)
g1qx1, g2hx2 = (288, 180, 0., 60.), (144, 90, 0., 120.)


def test_partition_exact_nesting():
    h = orc.Hntr(g4, g8)
    # the east-west walk starts one revolution early, so A indices come out in (imA, 2*imA] and are wrapped
    # by IA = 1 + (IAREV-1) % imA in the regrid loop (hntr.cpp:88, hntr.hpp:228)
    assert list(h.get("IMIN")) == [17, 19, 21, 23, 25, 27, 29, 31] and list(h.get("IMAX")) == [18, 20, 22, 24, 26, 28, 30, 32]
    assert list(h.get("JMIN")) == [1, 3, 5, 7] and list(h.get("JMAX")) == [2, 4, 6, 8]
    for n in ("FMIN", "FMAX", "GMIN", "GMAX"):
        assert np.all(h.get(n) == 0)
    assert np.all(h.get("JMAX") <= 8)


def test_partition_offset_wraps():
    """A half-cell offset makes the last B cell wrap around the date line (IMAX(imB) += imA, hntr.cpp:114)."""
    h = orc.Hntr((4, 2, 0.5, 5400.), (8, 4, 0.0, 2700.))
    assert list(h.get("IMIN")) == [10, 12, 14, 16] and list(h.get("IMAX")) == [11, 13, 15, 17]
    A = np.arange(32, dtype=float)
    B = h.regrid(None, A)
    a2 = A.reshape(4, 8)
    dA = h.get("dxypA")
    expect = (a2[0:2, [7, 0]] * dA[0:2, None]).sum() / (2 * dA[0:2].sum())     # wrapped cell (ib=4, jb=1)
    assert np.isclose(B[3], expect, rtol=1e-14)


@pytest.mark.parametrize("B,A", [(g2, g4), (g4, g2), (g1qx1, g2hx2), (g2hx2, g1qx1)])
def test_overlap_identities(B, A):
    h = orc.Hntr(B, A)
    for R in (1.0, 2.0):
        r, c, v = h.matrix(R)
        areaB = np.bincount(r, v, B[0] * B[1]); areaA = np.bincount(c, v, A[0] * A[1])
        np.testing.assert_allclose(areaB, np.repeat(h.get("dxypB"), B[0]) * R * R, rtol=1e-12)
        np.testing.assert_allclose(areaA, np.repeat(h.get("dxypA"), A[0]) * R * R, rtol=1e-12)
        assert abs(areaB.sum() / (4 * np.pi * R * R) - 1) < 1e-12
        assert np.all(v > areaB[r] * 1e-5)          # no spurious slivers
    r, c, v = h.matrix(None)                        # scaled_regrid_matrix: rows sum to 1
    np.testing.assert_allclose(np.bincount(r, v, B[0] * B[1]), 1.0, rtol=1e-13)


def test_regrid_conserves_and_matches_matrix():
    rng = np.random.default_rng(0)
    for B, A in ((g4, g8), (g1qx1, (720, 360, 0., 30.)), (g1qx1, (360, 180, 0., 60.)), (g2hx2, g1qx1)):
        h = orc.Hntr(B, A)
        src = rng.uniform(0, 1, A[0] * A[1])
        out = h.regrid(None, src)
        sa = (src.reshape(A[1], A[0]) * h.get("dxypA")[:, None]).sum()
        sb = (out.reshape(B[1], B[0]) * h.get("dxypB")[:, None]).sum()
        assert abs(sa / sb - 1) < 1e-12                      # test_hntr.cpp:308
        r, c, v = h.matrix(None)
        np.testing.assert_allclose(np.bincount(r, v * src[c], B[0] * B[1]), out, rtol=1e-12)   # :258-293


def test_regrid_vs_numpy_twin():
    rng = np.random.default_rng(1)
    for B, A in ((g2, g4), (g4, g8), (g8, g4), ((7, 5, 0.5, 36.0 * 60), (9, 11, 0.25, 180.0 * 60 / 11))):
        ho, ht = orc.Hntr(B, A, -99.0), oracle_np.Hntr(B, A, -99.0)
        for n in ("FMIN", "FMAX", "GMIN", "GMAX", "IMIN", "IMAX", "JMIN", "JMAX", "SINA", "SINB"):
            assert np.array_equal(ho.get(n), getattr(ht, n)), n
        src = rng.uniform(-1, 1, A[0] * A[1])
        W = (rng.uniform(0, 1, A[0] * A[1]) < 0.5).astype(float)
        for w in (None, W):
            for mp in (False, True):
                np.testing.assert_allclose(ho.regrid(w, src, mean_polar=mp), ht.regrid(w, src, mean_polar=mp), rtol=1e-13, atol=1e-15)


def test_datmis_and_polar_quirk():
    """Cells with zero total weight get DATMIS; one DATMIS cell in a polar row makes the whole row DATMIS (hntr.hpp:419-427)."""
    h = orc.Hntr(g4, g8, -7.0)
    src = np.ones(128); W = np.ones((8, 16)); W[0:2, 0:2] = 0
    out = h.regrid(W.ravel(), src).reshape(4, 8)
    assert out[0, 0] == -7.0 and np.all(out.ravel()[1:] == 1.0)
    outp = h.regrid(W.ravel(), src, mean_polar=True).reshape(4, 8)
    assert np.all(outp[0] == -7.0) and np.all(outp[3] == 1.0) and np.all(outp[1:3] == 1.0)
    i16 = (np.arange(128) % 7).astype(np.int16)
    assert np.array_equal(h.regrid(None, i16), h.regrid(None, i16.astype(float)))


def test_wtm_wtb():
    h = orc.Hntr(g4, g8)
    rng = np.random.default_rng(2)
    src, W = rng.uniform(0, 1, 128), rng.uniform(0, 1, 128)
    np.testing.assert_allclose(h.regrid(W, src, wtm=2.0, wtb=0.5), h.regrid(2.0 * W + 0.5, src), rtol=1e-14)


def test_spec_heuristic():
    with pytest.raises(orc.OracleError):
        orc.Hntr((144, 90, 0., 2.0), g1qx1)     # dlat in degrees, GridSpec.cpp:125-132


def test_golden_vectors():
    gold = np.load(os.path.join(GOLD, "hntr_kat.npz"))
    pairs = {"g8_to_g4": (g4, g8), "g1qx1_to_g2hx2": (g2hx2, g1qx1), "ghxh_to_g1qx1": (g1qx1, (720, 360, 0., 30.)),
             "g1x1_to_g1qx1": (g1qx1, (360, 180, 0., 60.))}
    for name, (B, A) in pairs.items():
        h = orc.Hntr(B, A)
        for t in ("FMIN", "FMAX", "GMIN", "GMAX", "IMIN", "IMAX", "JMIN", "JMAX"):
            assert np.array_equal(h.get(t), gold["%s_%s" % (name, t)]), (name, t)
        src = np.random.default_rng(5).uniform(0, 1, A[0] * A[1])
        np.testing.assert_allclose(h.regrid(None, src), gold[name + "_src_seed5_B"], rtol=1e-14)
