"""Per-warp cycle budget of smooth_group_kernel (needs tools/_trace/libicebin_b200_gtrace.so: build_variants.py gtrace
IB_GROUP_TRACE): waiting for a stage / math / stores, for the consumer warps, and waiting for a free stage, for the
producer -- medians over the CTAs of the LAST launch, in microseconds at the SM clock the run reports."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from icebin_b200 import api, synth
from icebin_b200._lib import lib
cfg = synth.config(3)
sh, em = cfg["sheets"][0], cfg["elevmask"][0]
g = api.GCMRegridder(cfg["gridA"], cfg["hcdefs"], True)
g.add_sheet(sh.name, sh)
M = g.regrid_matrices(sh.name, em, scale=True, correctA=True, sigma=(50e3, 50e3, 100.0)).matrix("IvE")
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
nvar = 512
X = torch.rand((M.shape[1], nvar), dtype=torch.float64, device="cuda")
Y = torch.empty((M.shape[0], nvar), dtype=torch.float64, device="cuda")
fn = lambda: M.apply_dev(X.data_ptr(), nvar, nvar, Y.data_ptr(), nvar, False, st.cuda_stream)
for _ in range(5):
    fn()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20):
    fn()
b.record(); torch.cuda.synchronize()
us = a.elapsed_time(b) / 20 * 1e3
nblk = 148
buf = np.zeros((nblk, 17, 8), np.uint64)
L = lib()
L.ib_debug_group_trace.argtypes = [C.c_void_p, C.c_int]
assert L.ib_debug_group_trace(buf.ctypes.data_as(C.c_void_p), nblk) == 0
t = buf.astype(np.float64)
np.save(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "group_trace.npy"), buf)
tot = np.median(t[:, :16, 3])
mhz = float(np.median(t[:, :16, 3] / np.maximum(t[:, :16, 5], 1.0))) * 1e3      # SM cycles per microsecond, from %globaltimer
out = {"us_per_launch": round(us, 1), "cta_cycles_median": tot, "sm_mhz": round(mhz), "cta_life_us": round(float(np.median(t[:, :16, 5])) / 1e3, 1)}
for k, name in ((0, "wait_full"), (4, "loads_check"), (1, "mma_and_rest"), (2, "store"), (3, "total")):
    out["consumer_" + name + "_us"] = [round(float(np.percentile(t[:, :16, k], p)) / mhz, 1) for p in (10, 50, 90)]
for k, name in enumerate(("wait_empty", "issue", "-", "total")):
    if name != "-":
        out["producer_" + name + "_us"] = [round(float(np.percentile(t[:, 16, k], p)) / mhz, 1) for p in (10, 50, 90)]
g0 = t[:, :16, 6].min()
out["start_us_min_med_max"] = [round(float(f(t[:, 0, 6] - g0)) / 1e3, 1) for f in (np.min, np.median, np.max)]
out["end_us_min_med_max"] = [round(float(f(t[:, :16, 7].max(axis=1) - g0)) / 1e3, 1) for f in (np.min, np.median, np.max)]
out["cta_life_us_min_med_max"] = [round(float(f(t[:, :16, 5].max(axis=1))) / 1e3, 1) for f in (np.min, np.median, np.max)]
print(json.dumps({k: v for k, v in out.items() if 'consumer' not in k and 'producer' not in k}))
print(json.dumps(out))
