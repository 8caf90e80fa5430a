"""Times the EvI and AvI applies (512 fields, Greenland 5 km) under the environment's SpMM knobs.  One JSON line.
Usage: IB_PAIR_CHUNK=64 python tools/evi_sweep.py [label]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from icebin_b200 import api, synth

PEAK = 6540.8
cfg = synth.config(3)
sh, em = cfg["sheets"][0], cfg["elevmask"][0]
g = api.GCMRegridder(cfg["gridA"], cfg["hcdefs"], True)
g.add_sheet(sh.name, sh)
rm = g.regrid_matrices(sh.name, em, scale=True, correctA=True)
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
S = st.cuda_stream
out = {"label": sys.argv[1] if len(sys.argv) > 1 else "", "env": {k: v for k, v in os.environ.items() if k.startswith("IB_")}}
nvar = 512
for spec in ("EvI", "AvI"):
    M = rm.matrix(spec)
    X = torch.rand((M.shape[1], nvar), dtype=torch.float64, device="cuda")
    Y = torch.empty((M.shape[0], nvar), dtype=torch.float64, device="cuda")
    fn = lambda: M.apply_dev(X.data_ptr(), nvar, nvar, Y.data_ptr(), nvar, True, S)
    for _ in range(5):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 200
    a.record()
    for _ in range(n):
        fn()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / n
    byt = 12 * M.nnz + 4 * (M.shape[0] + 1) + 8 * nvar * (M.shape[1] + M.shape[0])
    out[spec] = {"ms": round(ms, 5), "alg_GBs": round(byt / ms / 1e6, 1), "frac": round(byt / ms / 1e6 / PEAK, 4)}
print(json.dumps(out))
