"""Times the coupling-step applies (512 fields, Greenland 5 km): the stacked [EvI; AvI] one-pass apply, EvI alone and
AvI alone, back to back with CUDA events, for each value of the `stream_bps` (and `stream_spare`) knob.  One JSON line
per setting.  Usage: python tools/stream_sweep.py [bps ...] [--spare s ...]   (also: --once STACK|EvI|AvI for ncu)"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from icebin_b200 import api, synth

PEAK = 6540.8
args = sys.argv[1:]
once = None
if args and args[0] == "--once":
    once, args = args[1], args[2:]
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
nI, nE, nA = EvI.shape[1], EvI.shape[0], AvI.shape[0]
cases = {
    "STACK": (lambda: api.apply_many_dev([EvI, AvI], X.data_ptr(), nvar, nvar, [YE.data_ptr(), YA.data_ptr()], [nvar, nvar], True, S),
              12 * (EvI.nnz + AvI.nnz) + 8 * nvar * (nI + nE + nA), EvI.nnz + AvI.nnz),
    "EvI": (lambda: EvI.apply_dev(X.data_ptr(), nvar, nvar, YE.data_ptr(), nvar, True, S), 12 * EvI.nnz + 8 * nvar * (nI + nE), EvI.nnz),
    "AvI": (lambda: AvI.apply_dev(X.data_ptr(), nvar, nvar, YA.data_ptr(), nvar, True, S), 12 * AvI.nnz + 8 * nvar * (nI + nA), AvI.nnz),
}
if once:
    for _ in range(4):
        cases[once][0]()
    torch.cuda.synchronize()
    print("ok", once)
    sys.exit(0)
spares = [0]
if "--spare" in args:                       # e.g. `3 --spare 0 8 16 24`: CTA slots the one-pass kernel leaves free
    k = args.index("--spare")
    spares, args = [int(a) for a in args[k + 1:]], args[:k]
for bps, spare in [(b, sp) for b in ([int(a) for a in args] or [4]) for sp in spares]:
    api.tune("stream_bps", bps)
    api.tune("tma_bps", bps)
    api.tune("stream_spare", spare)
    out = {"bps": bps, "spare": spare}
    for name, (fn, byt, nnz) in cases.items():
        for _ in range(5):
            fn()
        best = 1e9
        for rep in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = 200
            a.record()
            for _ in range(n):
                fn()
            b.record(); torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b) / n)
        out[name] = {"us": round(1e3 * best, 2), "alg_GBs": round(byt / best / 1e6, 1), "frac": round(byt / best / 1e6 / PEAK, 4),
                     "Gnnz_s": round(nnz * nvar / best / 1e6, 1)}
    print(json.dumps(out), flush=True)
