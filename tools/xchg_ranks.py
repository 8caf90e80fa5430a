"""Multi-GPU check and timing of the peer-memory merge (csrc/xchg.cu) against the NCCL merge (multi.SheetMerge).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tools/xchg_ranks.py

Every rank owns one Greenland-5 km-shaped sheet (as bench.py --gpus N); both merges run on the same unscaled EvI / AvI
results and must agree to rounding; the peer merge must give identical bits on all owners of a cell.  Then both are
timed (CUDA events, max over ranks), eagerly and as a captured CUDA graph."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import bench
from icebin_b200 import api, multi
from icebin_b200._lib import check, lib

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
check(lib().ib_set_device(local))
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
st = torch.cuda.Stream(device=dev); torch.cuda.set_stream(st)
S = st.cuda_stream
nvar = 512
cfg = bench.make_config("greenland_5km", rank)
sh, em = cfg["sheets"][0], cfg["elevmask"][0]
gcm = api.GCMRegridder(cfg["gridA"], cfg["hcdefs"], True)
gcm.add_sheet(sh.name, sh)
rm = gcm.regrid_matrices(sh.name, em, scale=False, correctA=True)
EvI, AvI = rm.matrix("EvI"), rm.matrix("AvI")
X = torch.rand((EvI.shape[1], nvar), dtype=torch.float64, device=dev)
YE = torch.empty((EvI.shape[0], nvar), dtype=torch.float64, device=dev)
YA = torch.empty((AvI.shape[0], nvar), dtype=torch.float64, device=dev)


def spmm():
    api.apply_many_dev([EvI, AvI], X.data_ptr(), nvar, nvar, [YE.data_ptr(), YA.data_ptr()], [nvar, nvar], True, S)


nccl = multi.SheetMerge([(gcm.nE, EvI), (gcm.nA, AvI)], nvar, dev)
peer = multi.PeerMerge([(gcm.nE, EvI), (gcm.nA, AvI)], nvar, dev)
out = {"world": world, "shared_cells": peer.nshared}
# ---- agreement
spmm(); nccl.run([YE, YA], st); torch.cuda.synchronize()
refE, refA = YE.clone(), YA.clone()
for rep in range(4):
    spmm(); peer.merge([YE, YA], S); torch.cuda.synchronize()
    assert torch.allclose(YE, refE, rtol=1e-13, atol=0) and torch.allclose(YA, refA, rtol=1e-13, atol=0), "peer merge differs from the NCCL merge"
peer.status()
# identical bits on all owners: compare the shared rows through an all-gather
for p, (Y, n_sparse) in enumerate(((YE, gcm.nE), (YA, gcm.nA))):
    part = peer.parts[p]
    rows = torch.full((part["nshared"], nvar), float("nan"), dtype=torch.float64, device=dev)
    have = torch.from_numpy(part["loc"] >= 0).to(dev)
    rows[have] = Y[torch.from_numpy(part["loc"][part["loc"] >= 0]).long().to(dev)]
    allrows = [torch.empty_like(rows) for _ in range(world)]
    dist.all_gather(allrows, rows)
    for r in range(world):
        both = have & ~torch.isnan(allrows[r][:, 0])
        assert torch.equal(rows[both], allrows[r][both]), "owners of a shared cell hold different bits"
out["agree"] = True


def timed(fn, n=200):
    for _ in range(5):
        fn()
    dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / n], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()) * 1e3


# scaled mode: the sheets' SCALED matrices, only the shared cells travel
rms = gcm.regrid_matrices(sh.name, em, scale=True, correctA=True)
EvIs, AvIs = rms.matrix("EvI"), rms.matrix("AvI")
peer_s = multi.PeerMerge([(gcm.nE, EvIs), (gcm.nA, AvIs)], nvar, dev, scaled=True)
api.tune("stream_spare", 8)


def spmm_s():
    api.apply_many_dev([EvIs, AvIs], X.data_ptr(), nvar, nvar, [YE.data_ptr(), YA.data_ptr()], [nvar, nvar], True, S)


for rep in range(3):
    spmm_s(); peer_s.merge([YE, YA], S); torch.cuda.synchronize()
    assert torch.allclose(YE, refE, rtol=1e-12, atol=0) and torch.allclose(YA, refA, rtol=1e-12, atol=0), "scaled peer merge differs"
peer_s.status()
yps, lds = peer_s.args([YE, YA])
out["us_spmm_plus_scaled_peer_merge_eager"] = timed(lambda: (spmm_s(), peer_s.merge_raw(yps, lds, nvar, S)))
yp, ld = peer.args([YE, YA])
out["us_spmm_only"] = timed(spmm)
out["us_spmm_plus_nccl_merge_eager"] = timed(lambda: (spmm(), nccl.run([YE, YA], st)))
out["us_spmm_plus_peer_merge_eager"] = timed(lambda: (spmm(), peer.merge_raw(yp, ld, nvar, S)))
g = torch.cuda.CUDAGraph()
torch.cuda.synchronize(); dist.barrier()
with torch.cuda.graph(g, stream=st):
    for _ in range(8):
        spmm(); peer.merge_raw(yp, ld, nvar, S)
out["us_spmm_plus_peer_merge_graph"] = timed(g.replay, n=50) / 8
peer.status()
spmm(); peer.merge([YE, YA], S); torch.cuda.synchronize()
assert torch.allclose(YE, refE, rtol=1e-13, atol=0), "peer merge differs after the graph replays"
if rank == 0:
    print(json.dumps(out), flush=True)
dist.destroy_process_group()
