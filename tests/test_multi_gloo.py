"""CPU, world_size 2 over gloo: the host-side logic of the multi-sheet merge (icebin_b200/multi.py) -- which cells two
ranks must exchange -- plus the numpy statement of the merge against the oracle's VectorMultivec restatement."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from icebin_b200 import multi
from oracle import orc


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ids_per_rank, n_sparse, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        shared, loc, slot = multi.plan_shared(ids_per_rank[rank], n_sparse)
        # exchange of the shared rows, as SheetMerge.run does (here with torch CPU tensors)
        rng = np.random.default_rng(rank)
        nvar = 3
        num = rng.uniform(0, 1, (len(ids_per_rank[rank]), nvar))
        w = rng.uniform(1, 2, len(ids_per_rank[rank]))
        buf = torch.zeros((len(shared), nvar), dtype=torch.float64)
        wt = torch.zeros(len(shared), dtype=torch.float64)
        have = loc >= 0
        buf[torch.from_numpy(have)] = torch.from_numpy(num[loc[have]])
        wt[torch.from_numpy(have)] = torch.from_numpy(w[loc[have]])
        dist.all_reduce(buf)
        dist.all_reduce(wt)
        merged = num * (1.0 / w)[:, None]
        for r, k in enumerate(slot):
            if k >= 0:
                merged[r] = buf[k].numpy() * (1.0 / wt[k].item())
        out.put((rank, shared, loc, slot, num, w, merged))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_plan_and_merge_world2():
    n_sparse = 50
    ids = [np.array([7, 3, 11, 20, 21, 40], np.int64), np.array([21, 5, 3, 44, 40, 9, 8], np.int64)]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ids, n_sparse, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(2):
        r = q.get(timeout=120)
        res[r[0]] = r[1:]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for r in range(2):
        shared, loc, slot = res[r][0], res[r][1], res[r][2]
        assert list(shared) == [3, 21, 40]
        assert list(ids[r][loc]) == [3, 21, 40]
        assert sorted(np.flatnonzero(slot >= 0)) == sorted(loc)
    # merged values == the numpy statement == the oracle's VectorMultivec::to_dense
    contribs = [(ids[r], res[r][4], res[r][3]) for r in range(2)]
    ref = multi.merge_reference(contribs, n_sparse)
    idx = np.concatenate(ids); w = np.concatenate([res[0][4], res[1][4]]); v = np.concatenate([res[0][3], res[1][3]])
    with np.errstate(divide="ignore"):
        _, dense = orc.multivec_to_dense(idx, w, v, n_sparse, fill=np.nan)
    np.testing.assert_allclose(ref[~np.isnan(ref[:, 0])], dense.T[~np.isnan(dense[0])], rtol=1e-14)
    for r in range(2):
        np.testing.assert_allclose(res[r][5], ref[ids[r]], rtol=1e-14)


def test_plan_single_process():
    shared, loc, slot = multi.plan_shared([4, 2, 9], 12)
    assert len(shared) == 0 and len(loc) == 0 and list(slot) == [-1, -1, -1]
