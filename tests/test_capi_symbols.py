"""CPU: the C-ABI library loads, exports every symbol include/icebin_b200.h declares, and fails loudly (no CPU
fallback) when there is no CUDA device.  No compute calls are made here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "icebin_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(ib_[A-Za-z0-9_]+)\s*\(", txt)))


def test_library_builds_and_loads():
    from icebin_b200 import build
    so = build.build()
    assert os.path.exists(so)
    from icebin_b200 import _lib
    assert _lib.lib() is not None


def test_every_declared_symbol_is_exported_and_bound():
    from icebin_b200 import _lib
    L = C.CDLL(_lib.SO_PATH)
    syms = header_symbols()
    assert len(syms) >= 40
    for s in syms:
        assert hasattr(L, s), "declared in include/icebin_b200.h but not exported: " + s
    assert set(syms) == set(_lib.SIGNATURES), set(syms) ^ set(_lib.SIGNATURES)


def test_header_cites_reference():
    txt = open(os.path.join(ROOT, "include", "icebin_b200.h")).read()
    for cite in ("GCMRegridder.hpp:346-352", "RegridMatrices_Dynamic.cpp:334-402", "hntr.hpp:385-435", "multivec.cpp:35-81",
                 "icebin_cython.cpp:159-192"):
        assert cite in txt


def test_no_cpu_fallback():
    """Without a GPU every compute entry point must return IB_ERR_CUDA; with one this test is a no-op."""
    from icebin_b200 import api
    if api.device_count() > 0:
        pytest.skip("CUDA device present")
    from icebin_b200 import synth
    gA, sh = synth.make_gcm_and_sheet("greenland_50km")
    with pytest.raises(api.IceBinError) as e:
        api.GCMRegridder(gA, synth.hcdefs_example1(), True)
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)
    h = api.Hntr(17.17, api.GRIDS["g2hx2"], api.GRIDS["g1qx1"])       # host-only: partition tables
    assert h.get("IMAX")[-1] == 288 + 288 and len(h.get("SINA")) == 181
    with pytest.raises(api.IceBinError) as e:
        h.regrid(None, np.zeros(288 * 180))
    assert e.value.code == -2
    # the entry points either side of the path (exchange-grid generation, variable transformation, device buffers)
    import scipy.sparse
    from icebin_b200._lib import lib, vp
    with pytest.raises(api.IceBinError) as e:
        api.exgrid_rectilinear([np.array([[0., 0.], [1., 0.], [1., 1.]])], [0], np.arange(3.), np.arange(3.))
    assert e.value.code == -2
    with pytest.raises(api.IceBinError) as e:
        api.vartrans_dev(0x1000, 2, 1, scipy.sparse.identity(2, format="csc"), np.zeros(2), 0x2000, 2)
    assert e.value.code == -2
    p = vp()
    assert lib().ib_dev_alloc(C.byref(p), 64) == -2 and not p.value


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under icebin_b200/ may import, link or load it."""
    pat = re.compile(r"(from\s+oracle|import\s+oracle|liboracle|oracle[/\\.](orc|oracle_np|liboracle)|#include\s*[<\"].*oracle)")
    for dp, _, fs in os.walk(os.path.join(ROOT, "icebin_b200")):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dp, f), errors="replace").read()
                assert not pat.search(src), f


def test_host_side_helpers():
    """coo_matvec (icebin_cython.cpp:159-192) is the one entry point computed on the host: a single-vector convenience of
    the Python binding, documented as such in the header.  The VectorMultivec merge runs on the device."""
    import torch
    from icebin_b200 import api
    yy = np.full(4, 7.0)
    xx = np.array([1.0, np.nan, 2.0])
    api.coo_matvec(yy, xx, True, 4, 3, [0, 0, 2, 3], [0, 1, 2, 1], [1.0, 5.0, 3.0, 4.0])
    assert list(yy) == [1.0, 7.0, 6.0, 7.0]          # rows 1 and 3 untouched, NaN input skipped
    if not torch.cuda.is_available():
        idx = np.array([3, 1, 3, 0]); w = np.array([2.0, 1.0, 6.0, 4.0]); v = np.array([[4.0, 1], [5.0, 2], [12.0, 3], [8.0, 4]])
        with pytest.raises(api.IceBinError) as e:
            api.multivec_to_dense(idx, w, v, 5, fill=-1.0)
        assert e.value.code == -2


@pytest.mark.gpu
def test_multivec_to_dense_bit_exact():
    """VectorMultivec::to_dense_scale + to_dense (multivec.cpp:35-81) on the device: entries sorted by cell (stable), summed in
    input order with mul-then-add -- the bits of the reference's sequential loop."""
    from icebin_b200 import api
    from oracle import orc
    idx = np.array([3, 1, 3, 0]); w = np.array([2.0, 1.0, 6.0, 4.0]); v = np.array([[4.0, 1], [5.0, 2], [12.0, 3], [8.0, 4]])
    with np.errstate(divide="ignore"):
        s1, d1 = api.multivec_to_dense(idx, w, v, 5, fill=-1.0)
        s2, d2 = orc.multivec_to_dense(idx, w, v, 5, fill=-1.0)
    assert np.array_equal(s1, s2) and np.array_equal(d1, d2)
    rng = np.random.default_rng(0)
    n, nE, nvar = 200000, 5000, 3
    idx = rng.integers(0, nE - 100, n); w = rng.uniform(0.1, 1, n); v = rng.uniform(-1, 1, (n, nvar))
    with np.errstate(divide="ignore"):
        s1, d1 = api.multivec_to_dense(idx, w, v, nE)
        s2, d2 = orc.multivec_to_dense(idx, w, v, nE)
    assert np.array_equal(s1, s2) and np.array_equal(d1, d2, equal_nan=True)
    with pytest.raises(api.IceBinError):
        api.multivec_to_dense(np.array([0, 7]), np.ones(2), np.ones((2, 1)), 5)
