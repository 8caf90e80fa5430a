"""Per-matrix build times (wall clock, device-resident elevmask) for one synthetic sheet.
Usage: python tools/build_time.py greenland_1km|antarctica_1km|greenland_5km [reps]"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from icebin_b200 import api, synth
shape = sys.argv[1]; reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
hc = synth.hcdefs_conserv(40)
gA = bench.gridA_global()
anchor = (40, 2) if shape.startswith("antarctica") else (40, 60)
_, sh = synth.make_gcm_and_sheet(shape, anchor=anchor, gridA=gA, name=shape)
em = torch.from_numpy(synth.make_elevmask(sh, hc[-1], 0)).cuda()
g = api.GCMRegridder(gA, hc, True); g.add_sheet(sh.name, sh)
out = {}
for it in range(reps):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    rm = g.regrid_matrices_dev(sh.name, em.data_ptr(), scale=True, correctA=True)
    torch.cuda.synchronize(); t = {"rm": 1e3 * (time.perf_counter() - t0)}
    for s in bench.SIX:
        t1 = time.perf_counter(); l0 = api.launch_count()
        M = rm.matrix(s); torch.cuda.synchronize()
        t[s] = round(1e3 * (time.perf_counter() - t1), 3); t[s + "_launches"] = api.launch_count() - l0
        del M
    t["total"] = round(1e3 * (time.perf_counter() - t0), 3)
    out[it] = t
print(json.dumps(out[reps - 1])); print(json.dumps({k: out[1][k] for k in out[1] if not k.endswith("_launches")}))
