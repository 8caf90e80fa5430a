"""A small tour of every kernel family for compute-sanitizer (config 1 sizes)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from icebin_b200 import api, synth
cfg = synth.config(1)
gA, sh, hc, em = cfg["gridA"], cfg["sheets"][0], cfg["hcdefs"], cfg["elevmask"][0]
g = api.GCMRegridder(gA, hc, True)
g.add_sheet(sh.name, sh)
rm = g.regrid_matrices(sh.name, em, scale=True, correctA=True)
mats = {s: rm.matrix(s) for s in ("EvI", "AvI", "IvE", "IvA", "EvA", "AvE", "EvX", "XvA")}
for nvar in (3, 64, 512):
    X = np.random.default_rng(0).uniform(-1, 1, (mats["EvI"].shape[1], nvar))
    YE, YA = api.apply_many([mats["EvI"], mats["AvI"]], X)
    mats["IvE"].apply(YE, layout="cell")
    mats["IvA"].apply(YA, layout="cell")
Ws = g.regrid_matrices(sh.name, em, scale=True, correctA=True, sigma=(100e3, 100e3, 300.0)).matrix("IvE")
Ws.apply(np.ones((2, Ws.shape[1])))
dI, dE = api.SparseSet(identity=sh.nI), api.SparseSet()
rm.matrix_d("EvI", (dE, dI), scale=False, correctA=False)
h = api.Hntr(17.17, api.GRIDS["g2hx2"], api.GRIDS["g1qx1"])
A = np.random.default_rng(1).uniform(0, 1, 288 * 180)
h.regrid(None, A, mean_polar=True); h.regrid(A, A); h.regrid(None, (A * 100).astype(np.int16)); h.overlap(1.0)
print("sanitize tour ok")
