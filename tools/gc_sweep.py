"""Times the I-side applies (IvE, IvA; 512 fields, Greenland 5 km) under the SpMM variant selected by the environment
(IB_SPMM_NOGC=1: previous kernels; IB_GC_BPS=2|3: distinct-column group kernel at 2 or 3 CTAs/SM).  One JSON line.
Usage: IB_GC_BPS=3 python tools/gc_sweep.py [label]"""
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
for spec in ("IvE", "IvA"):
    M = rm.matrix(spec)
    for nvar in (512, 64):
        X = torch.rand((M.shape[1], nvar), dtype=torch.float64, device="cuda")
        Y = torch.empty((M.shape[0], nvar), dtype=torch.float64, device="cuda")
        fn = lambda: M.apply_dev(X.data_ptr(), nvar, nvar, Y.data_ptr(), nvar, True, S)
        for _ in range(5):
            fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 100
        a.record()
        for _ in range(n):
            fn()
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / n
        byt = 12 * M.nnz + 4 * (M.shape[0] + 1) + 8 * nvar * (M.shape[1] + M.shape[0])
        out["%s_%d" % (spec, nvar)] = {"ms": round(ms, 5), "alg_GBs": round(byt / ms / 1e6, 1), "frac": round(byt / ms / 1e6 / PEAK, 4)}
print(json.dumps(out))
