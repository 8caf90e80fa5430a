"""Runs a few applies of one matrix of config 3 (for ncu).  Usage: python tools/apply_one.py IvE [sigma]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from icebin_b200 import api, synth
spec = sys.argv[1]
sigma = (50e3, 50e3, 100.0) if len(sys.argv) > 2 else (0, 0, 0)
cfg = synth.config(3)
sh, em = cfg["sheets"][0], cfg["elevmask"][0]
g = api.GCMRegridder(cfg["gridA"], cfg["hcdefs"], True)
g.add_sheet(sh.name, sh)
M = g.regrid_matrices(sh.name, em, scale=True, correctA=True, sigma=sigma).matrix(spec)
nvar = 512
X = torch.rand((M.shape[1], nvar), dtype=torch.float64, device="cuda")
Y = torch.empty((M.shape[0], nvar), dtype=torch.float64, device="cuda")
for _ in range(4):
    M.apply_dev(X.data_ptr(), nvar, nvar, Y.data_ptr(), nvar)
torch.cuda.synchronize()
print("ok", M.shape, M.nnz)
