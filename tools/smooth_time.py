"""Times the sigma-smoothed IvE apply (configs[2]) at 512 fields: register-blocked group kernel vs the previous kernels."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from icebin_b200 import api, synth
cfg = synth.config(3)
sh, em = cfg["sheets"][0], cfg["elevmask"][0]
g = api.GCMRegridder(cfg["gridA"], cfg["hcdefs"], True)
g.add_sheet(sh.name, sh)
M = g.regrid_matrices(sh.name, em, scale=True, correctA=True, sigma=(50e3, 50e3, 100.0)).matrix("IvE")
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
nvar = 512
X = torch.rand((M.shape[1], nvar), dtype=torch.float64, device="cuda")
Y = torch.empty((M.shape[0], nvar), dtype=torch.float64, device="cuda")
alg = 12 * M.nnz + 4 * (M.shape[0] + 1) + 8 * nvar * (M.shape[0] + M.shape[1])
out = {"nnz": M.nnz, "shape": list(M.shape)}
ref = None
for label, knob in (("group", 1), ("previous", 0)):
    api.tune("use_group", knob)
    for fc in (False, True):
        fn = lambda: M.apply_dev(X.data_ptr(), nvar, nvar, Y.data_ptr(), nvar, fc, st.cuda_stream)
        for _ in range(3):
            fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(50):
            fn()
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 50
        out["%s fc=%d" % (label, fc)] = {"us": round(1e3 * ms, 1), "alg_GBs": round(alg / ms / 1e6), "frac": round(alg / ms / 1e6 / 6540.8, 3)}
        if fc:
            if ref is None:
                ref = Y.clone()
            else:
                out["max_rel_diff_between_kernels"] = float(((Y - ref).abs() / ref.abs().clamp_min(1e-300)).max())
api.tune("use_group", 1)
print(json.dumps(out))
