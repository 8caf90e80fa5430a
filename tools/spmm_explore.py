"""Times apply_dev on synthetic CSR patterns (streaming vs random gathers) to separate kernel limits from DRAM
access-pattern limits.  Usage: python tools/spmm_explore.py [nvar]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from icebin_b200 import api

nvar = int(sys.argv[1]) if len(sys.argv) > 1 else 512
dev = torch.device("cuda")
st = torch.cuda.Stream(); torch.cuda.set_stream(st)


def timeit(W, X, Y, n=50):
    for _ in range(5):
        W.apply_dev(X.data_ptr(), nvar, nvar, Y.data_ptr(), nvar, True, st.cuda_stream)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        W.apply_dev(X.data_ptr(), nvar, nvar, Y.data_ptr(), nvar, True, st.cuda_stream)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def case(name, nrow, ncol, rowlen, cols):
    rowptr = np.arange(nrow + 1, dtype=np.int32) * rowlen
    val = np.random.default_rng(0).uniform(0, 1, nrow * rowlen)
    W = api.Weighted.from_csr(nrow, ncol, rowptr, cols.astype(np.int32), val)
    X = torch.rand((ncol, nvar), dtype=torch.float64, device=dev)
    Y = torch.empty((nrow, nvar), dtype=torch.float64, device=dev)
    ms = timeit(W, X, Y)
    touched = len(np.unique(cols))
    byt = 12 * nrow * rowlen + 8 * nvar * (touched + nrow)
    print("%-34s rows=%6d len=%4d  %.4f ms  %.0f GB/s (alg)" % (name, nrow, rowlen, ms, byt / ms / 1e6), flush=True)


ncol = 69000
rng = np.random.default_rng(1)
for rowlen in (2, 32, 128, 690):
    nrow = ncol // rowlen
    seq = np.arange(nrow * rowlen)
    case("sequential", nrow, ncol, rowlen, seq)
    perm = rng.permutation(ncol)[: nrow * rowlen]
    case("random perm (sorted within row)", nrow, ncol, rowlen, np.sort(perm.reshape(nrow, rowlen), axis=1).ravel())
# write-dominated: IvE-like, 2 nnz per row, few columns
nrow = 69000
cols = np.sort(rng.integers(0, 1120, (nrow, 2)), axis=1).ravel()
case("IvE-like (2/row, 1120 cols)", nrow, 1120, 2, cols)
