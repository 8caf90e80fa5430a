"""CPU check of the distinct-column group plan (icebin_b200/csrc/gc_plan.cuh, the routine the device runs) and of the
per-warp bookkeeping of spmm_gc_kernel, replayed on the host by tests/cpp/gc_emul.cpp, against the oracle's apply."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import make_oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def emul_lib():
    src = os.path.join(ROOT, "tests", "cpp", "gc_emul.cpp")
    hdr = os.path.join(ROOT, "icebin_b200", "csrc", "gc_plan.cuh")
    so = os.path.join(ROOT, "tests", "cpp", "libgc_emul.so")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-x", "c++", src, "-o", so])
    L = C.CDLL(so)
    L.gc_emul.argtypes = [C.c_void_p] * 3 + [C.c_int64, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    return L


def csr_of(W):
    order = np.lexsort((W.col, W.row))
    rowptr = np.zeros(W.shape[0] + 1, np.int32)
    np.cumsum(np.bincount(W.row, minlength=W.shape[0]), out=rowptr[1:])
    return rowptr, np.ascontiguousarray(W.col[order], np.int32), np.ascontiguousarray(W.val[order])


def run(L, rowptr, col, val, X):
    nrow = len(rowptr) - 1
    Y = np.full((nrow, X.shape[1]), np.nan)
    stats = np.zeros(3, np.int64)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = L.gc_emul(p(rowptr), p(col), p(val), nrow, p(X), X.shape[1], p(Y), p(stats))
    assert rc == 0, "invariant %d violated" % rc
    return Y, stats


@pytest.mark.parametrize("spec", ["IvE", "IvA", "EvA"])
def test_gc_plan_matches_oracle_cfg2(cfg2, spec):
    go = make_oracle(cfg2)
    Wo = go.matrix(0, cfg2["elevmask"][0], spec, scale=True, correctA=True)
    rowptr, col, val = csr_of(Wo)
    assert np.diff(rowptr).max() <= 8
    X = np.random.default_rng(1).uniform(-1, 1, (Wo.shape[1], 3))
    Y, stats = run(emul_lib(), rowptr, col, val, X)
    assert np.array_equal(Y.T, Wo.apply(np.ascontiguousarray(X.T)))           # bit-identical: rows are never split
    if spec in ("IvE", "IvA"):
        assert stats[1] < 0.25 * Wo.nnz       # the point of the plan: few distinct X rows per group


def test_gc_plan_random_short_rows():
    """Adversarial patterns: empty rows, rows of exactly 8 distinct columns, no sharing at all, ragged tail."""
    rng = np.random.default_rng(5)
    L = emul_lib()
    for nrow, ncol, maxlen in ((1, 5, 3), (31, 9, 8), (32, 4000, 8), (33, 10, 8), (1000, 17, 8), (777, 100000, 8), (64, 3, 0)):
        lens = rng.integers(0, maxlen + 1, nrow)
        rowptr = np.zeros(nrow + 1, np.int32)
        np.cumsum(lens, out=rowptr[1:])
        col = np.concatenate([np.sort(rng.choice(ncol, min(n, ncol), replace=False)) for n in lens] + [np.zeros(0, np.int64)]).astype(np.int32)
        rowptr = np.zeros(nrow + 1, np.int32)
        np.cumsum([min(n, ncol) for n in lens], out=rowptr[1:])
        val = rng.uniform(-1, 1, len(col))
        X = rng.uniform(-1, 1, (ncol, 2))
        Y, _ = run(L, rowptr, col, val, X)
        ref = np.zeros((nrow, 2))
        for r in range(nrow):
            for j in range(rowptr[r], rowptr[r + 1]):
                ref[r] = ref[r] + val[j] * X[col[j]]
        assert np.array_equal(Y, ref)
