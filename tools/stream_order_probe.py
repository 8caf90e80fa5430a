"""How much of the one-pass apply's time is DRAM access ORDER?  Synthetic wide matrices (2 nonzeros per column, groups of
~61 columns feeding one pair of rows, like EvI at 5 km) whose sorted record stream reads X rows (4 KB each at 512 fields)
  seq    : in ascending order (pure streaming),
  runN   : in runs of N adjacent rows, runs in random order,
  rand   : in random order.
Same plan shape (records, events) in every case; only the order of the row reads differs.  One JSON line per case."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import scipy.sparse as sp
import torch
from icebin_b200 import api

PEAK = 6540.8
ncol, nvar, gsz = 68763, 512, 61
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
S = st.cuda_stream
X = torch.rand((ncol, nvar), dtype=torch.float64, device="cuda")
rng = np.random.default_rng(0)


def case(name, order):
    """order[i] = column read at position i of the stream"""
    ngroup = -(-ncol // gsz)
    grp = np.empty(ncol, np.int64)
    grp[order] = np.arange(ncol) // gsz          # columns of one group are consecutive in the stream
    rows = np.concatenate([2 * grp, 2 * grp + 1])
    cols = np.concatenate([np.arange(ncol), np.arange(ncol)])
    M = sp.csr_matrix((rng.uniform(0.1, 1, 2 * ncol), (rows, cols)), shape=(2 * ngroup, ncol))
    W = api.Weighted.from_csr(M.shape[0], M.shape[1], M.indptr, M.indices, M.data)
    Y = torch.empty((M.shape[0], nvar), dtype=torch.float64, device="cuda")
    fn = lambda: W.apply_dev(X.data_ptr(), nvar, nvar, Y.data_ptr(), nvar, True, S)
    for _ in range(5):
        fn()
    best = 1e9
    for rep in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(100):
            fn()
        b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / 100)
    byt = 12 * M.nnz + 8 * nvar * (ncol + M.shape[0])
    ref = M @ X[:, :4].cpu().numpy()
    err = float(np.abs(Y[:, :4].cpu().numpy() - ref).max())
    print(json.dumps({"case": name, "us": round(1e3 * best, 2), "alg_GBs": round(byt / best / 1e6, 1), "frac": round(byt / best / 1e6 / PEAK, 4), "err": err}), flush=True)


case("seq", np.arange(ncol))
for n in (16, 4, 2):
    nb = -(-ncol // n)
    blocks = rng.permutation(nb)
    order = (blocks[:, None] * n + np.arange(n)[None, :]).ravel()
    order = order[order < ncol]
    case("run%d" % n, order)
case("rand", rng.permutation(ncol))
