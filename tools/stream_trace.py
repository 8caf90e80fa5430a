"""Per-CTA timeline of spmm_stream_kernel (needs the instrumented library, see tools/build_trace.py).
Runs the stacked [EvI; AvI] apply (or EvI alone: argument EvI) a few times back to back and prints, for the LAST launch,
the distribution over CTAs of: start, producer released (previous kernel complete), first stage landed, streaming done,
flush done -- in microseconds relative to the earliest CTA start."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from icebin_b200 import api, synth
from icebin_b200._lib import lib

which = sys.argv[1] if len(sys.argv) > 1 else "STACK"
cfg = synth.config(3)
sh, em = cfg["sheets"][0], cfg["elevmask"][0]
g = api.GCMRegridder(cfg["gridA"], cfg["hcdefs"], True)
g.add_sheet(sh.name, sh)
rm = g.regrid_matrices(sh.name, em, scale=True, correctA=True)
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
S = st.cuda_stream
nvar = 512
EvI, AvI = rm.matrix("EvI"), rm.matrix("AvI")
X = torch.rand((EvI.shape[1], nvar), dtype=torch.float64, device="cuda")
YE = torch.empty((EvI.shape[0], nvar), dtype=torch.float64, device="cuda")
YA = torch.empty((AvI.shape[0], nvar), dtype=torch.float64, device="cuda")
if which == "STACK":
    fn = lambda: api.apply_many_dev([EvI, AvI], X.data_ptr(), nvar, nvar, [YE.data_ptr(), YA.data_ptr()], [nvar, nvar], True, S)
else:
    fn = lambda: EvI.apply_dev(X.data_ptr(), nvar, nvar, YE.data_ptr(), nvar, True, S)
for _ in range(10):
    fn()
torch.cuda.synchronize()
nblk = int(sys.argv[2]) if len(sys.argv) > 2 else 444
buf = np.zeros((nblk, 8), np.uint64)
L = lib()
L.ib_debug_stream_trace.argtypes = [C.c_void_p, C.c_int]
rc = L.ib_debug_stream_trace(buf.ctypes.data_as(C.c_void_p), nblk)
assert rc == 0, rc
np.save(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "trace_%s.npy" % which), buf)
t = buf.astype(np.float64)
t0 = t[:, 0].min()
t = (t - t0) / 1e3
names = ["start", "producer released", "first stage landed", "streaming done", "flush done", "producer done"]
out = {"which": which, "nblk": nblk}
for k, nm in enumerate(names):
    c = t[:, k]
    out[nm] = {"min": round(float(c.min()), 2), "p10": round(float(np.percentile(c, 10)), 2), "med": round(float(np.median(c)), 2),
               "p90": round(float(np.percentile(c, 90)), 2), "max": round(float(c.max()), 2)}
out["stream phase per CTA (first landed -> done)"] = {"med": round(float(np.median(t[:, 3] - t[:, 2])), 2), "max": round(float((t[:, 3] - t[:, 2]).max()), 2)}
out["flush per CTA"] = {"med": round(float(np.median(t[:, 4] - t[:, 3])), 2), "max": round(float((t[:, 4] - t[:, 3]).max()), 2)}
print(json.dumps(out))
