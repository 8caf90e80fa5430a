"""Builds the six matrices (+ optionally a smoothed IvE) for one sheet shape twice (for ncu: profile the 2nd pass).
Usage: python tools/build_one.py greenland_1km [smooth]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from icebin_b200 import api, synth
shape = sys.argv[1] if len(sys.argv) > 1 else "greenland_1km"
hc = synth.hcdefs_conserv(40)
gA, sh = synth.make_gcm_and_sheet(shape, anchor=(40, 60) if "greenland" in shape else (40, 2))
em = synth.make_elevmask(sh, hc[-1], 0)
g = api.GCMRegridder(gA, hc, True)
g.add_sheet(sh.name, sh)
for rep in range(2):
    l0 = api.launch_count()
    t0 = time.perf_counter()
    rm = g.regrid_matrices(sh.name, em, scale=True, correctA=True)
    mats = [rm.matrix(s) for s in ("EvI", "AvI", "IvE", "IvA", "EvA", "AvE")]
    torch.cuda.synchronize()
    print("pass", rep, "launches", api.launch_count() - l0, "ms", 1e3 * (time.perf_counter() - t0), flush=True)
    del mats
if len(sys.argv) > 2:
    M = g.regrid_matrices(sh.name, em, scale=True, correctA=True, sigma=(50e3, 50e3, 100.0)).matrix("IvE")
    print("smoothed nnz", M.nnz)
