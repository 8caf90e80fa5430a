"""Strong scaling of ONE ice sheet's six-matrix build over N GPUs (not yet measured: written at the end of round 1).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29551 \
        tools/bench_build_sharded.py [--shape greenland_1km|antarctica_1km] [--reps 7]

Every rank generates the same synthetic sheet, keeps the GCM cells multi.partition_cellsA gives it
(GCMRegridder_Standard::filter_cellsA, GCMRegridder.cpp:154-164; the reference shards global_ec the same way,
modele/global_ec.cpp:818-893), and builds {EvI, AvI, IvE, IvA, EvA, AvE} (unscaled, correctA) for its domain; the job's build
time is the slowest rank's median.  AvI / EvI rows of the domains are disjoint (tests/test_filter_cellsA.py), so the
result stays distributed; ib_w_combine_chunks would gather it.  Rank 0 prints one JSON line.
"""
import argparse, json, os, statistics, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from icebin_b200 import api, multi, synth
from icebin_b200._lib import check, lib

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="greenland_1km")
ap.add_argument("--reps", type=int, default=7)
args = ap.parse_args()
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
check(lib().ib_set_device(local))
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)

hc = synth.hcdefs_conserv(40)
im, jm = 144, 90
ids = np.arange(im * jm, dtype=np.int64)
full = synth.GridA(im, jm, im * jm, ids, (synth.A_W * synth.A_H) * (1.0 + 0.08 * np.sin(0.37 * (ids % im) + 0.21 * (ids // im))))
anchor = (40, 2) if args.shape.startswith("antarctica") else (40, 60)
_, sh = synth.make_gcm_and_sheet(args.shape, anchor=anchor, gridA=full, name=args.shape)
em = synth.make_elevmask(sh, hc[-1], 0)
g = api.GCMRegridder(full, hc, True)
g.add_sheet(sh.name, sh)
nX_full = g.nX(sh.name)
if world > 1:
    keep = multi.partition_cellsA(sh.ex_ij[:, 0], full.nA, world)
    g.filter_cellsA(keep[rank])
nX_mine = g.nX(sh.name)
SPECS = ("EvI", "AvI", "IvE", "IvA", "EvA", "AvE")
ts, nnz = [], {}
for it in range(args.reps + 3):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    rm = g.regrid_matrices(sh.name, em, scale=False, correctA=True)
    mats = {s: rm.matrix(s) for s in SPECS}
    torch.cuda.synchronize()
    if it >= 3:
        ts.append(1e3 * (time.perf_counter() - t0))
    nnz = {s: m.nnz for s, m in mats.items()}
med = statistics.median(ts)
t = torch.tensor([med, float(nX_mine)], dtype=torch.float64, device=dev)
tmax = t.clone()
if world > 1:
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"what": "six-matrix build of one sheet, GCM-cell domains over the GPUs (strong scaling)", "shape": args.shape,
                      "n_gpus": world, "nX": nX_full, "nX_largest_domain": int(tmax[1].item()), "build_ms_slowest_rank_median": float(tmax[0].item()),
                      "build_ms_rank0": ts, "nnz_rank0": nnz, "scaling": "strong"}), flush=True)
if world > 1:
    dist.destroy_process_group()
