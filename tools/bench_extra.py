"""Secondary measurements (BASELINE.md section 4 rows other than the headline): Hntr at ETOPO1 scale, full matrix
builds at 5 km and 1 km, and the I-side applies (IvE plain / smoothed, IvA).  Prints one JSON object.
Usage: python tools/bench_extra.py [--skip-ais] [--out profiles/xxx.json]"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from icebin_b200 import api, synth

ap = argparse.ArgumentParser()
ap.add_argument("--skip-ais", action="store_true")
ap.add_argument("--skip-1km", action="store_true")
ap.add_argument("--out", default=None)
ap.add_argument("--only", default="")
args = ap.parse_args()
PEAK = 6540.8
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
dev = torch.device("cuda")
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
S = st.cuda_stream
res = {"hbm_peak_gbs": PEAK}


def ev_time(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n


# ---------------------------------------------------------------- Hntr, config 4
G = api.GRIDS
A = G["g1mx1m"]
src16 = torch.randint(-500, 4000, (A.size,), dtype=torch.int16, device=dev)
src64 = torch.rand(A.size, dtype=torch.float64, device=dev)
w64 = (torch.rand(A.size, device=dev) < 0.3).to(torch.float64)
res["hntr"] = {}
for bname in ("g2hx2", "g1x1", "ghxh", "g1qx1"):
    if args.only and args.only != "hntr":
        break
    B = G[bname]
    h = api.Hntr(17.17, B, A, 0.0)
    out = torch.empty(B.size, dtype=torch.float64, device=dev)
    for tag, wt, src, dt, byt in (("i16_const", 0, src16, 1, 2 * A.size + 8 * B.size), ("f64_f64", w64.data_ptr(), src64, 0, 16 * A.size + 8 * B.size)):
        ms = ev_time(lambda: h.regrid_dev(wt, src.data_ptr(), dt, out.data_ptr(), True, 1.0, 0.0, S))
        res["hntr"]["g1mx1m->%s %s" % (bname, tag)] = {"ms": ms, "alg_GBs": byt / ms / 1e6, "frac_of_measured_peak": byt / ms / 1e6 / PEAK}
        print("hntr", bname, tag, "%.3f ms  %.0f GB/s" % (ms, byt / ms / 1e6), flush=True)
del src16, src64, w64


# ---------------------------------------------------------------- builds
def build_case(name, cfg, specs=("EvI", "AvI", "IvE", "IvA", "EvA", "AvE"), reps=5):
    gcm = api.GCMRegridder(cfg["gridA"], cfg["hcdefs"], True)
    out = {}
    for sh, em in zip(cfg["sheets"], cfg["elevmask"]):
        gcm.add_sheet(sh.name, sh)
        t = []
        nnz = {}
        for it in range(reps + 2):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            rm = gcm.regrid_matrices(sh.name, em, scale=True, correctA=True)
            mats = {s: rm.matrix(s) for s in specs}
            torch.cuda.synchronize()
            if it >= 2:
                t.append(1e3 * (time.perf_counter() - t0))
            nnz = {s: m.nnz for s, m in mats.items()}
            del mats, rm
        out[sh.name] = {"nX": sh.nX, "nI": sh.nI, "six_matrix_build_ms_median": float(np.median(t)), "min": float(min(t)), "nnz": nnz}
        print("build", name, sh.name, out[sh.name], flush=True)
    return out


if args.only == "hntr":
    print(json.dumps(res)); sys.exit(0)
res["build"] = {}
if args.only != "smooth":
    res["build"]["cfg2_greenland_5km"] = build_case("cfg2", synth.config(2))
if not args.skip_1km and args.only != "smooth":
    hc = synth.hcdefs_conserv(40)
    im, jm = 144, 90
    ids = np.arange(im * jm, dtype=np.int64)
    full = synth.GridA(im, jm, im * jm, ids, (synth.A_W * synth.A_H) * (1.0 + 0.08 * np.sin(0.37 * (ids % im) + 0.21 * (ids // im))))
    _, gr = synth.make_gcm_and_sheet("greenland_1km", anchor=(40, 60), gridA=full, name="greenland_1km")
    res["build"]["cfg5_greenland_1km"] = build_case("cfg5", dict(gridA=full, sheets=[gr], hcdefs=hc, elevmask=[synth.make_elevmask(gr, hc[-1], 0)]))
    del gr
    if not args.skip_ais:
        _, an = synth.make_gcm_and_sheet("antarctica_1km", anchor=(40, 2), gridA=full, name="antarctica_1km")
        res["build"]["cfg5_antarctica_1km"] = build_case("cfg5", dict(gridA=full, sheets=[an], hcdefs=hc, elevmask=[synth.make_elevmask(an, hc[-1], 1)]), reps=3)
        del an

# ---------------------------------------------------------------- I-side applies, config 3
cfg = synth.config(3)
sh, em = cfg["sheets"][0], cfg["elevmask"][0]
gcm = api.GCMRegridder(cfg["gridA"], cfg["hcdefs"], True)
gcm.add_sheet(sh.name, sh)
nvar = 512
res["apply"] = {}
ts = []
for it in range(4):          # first pass warms the allocator pool and loads the kernels
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    rms = gcm.regrid_matrices(sh.name, em, scale=True, correctA=True, sigma=(50e3, 50e3, 100.0))
    IvEs = rms.matrix("IvE")
    torch.cuda.synchronize()
    ts.append(1e3 * (time.perf_counter() - t0))
res["build"]["smoothed_IvE_5km_sigma50km_ms"] = {"median_of_3_after_warmup": float(np.median(ts[1:])), "first_call": ts[0], "nnz": IvEs.nnz}
print("smoothed IvE build", res["build"]["smoothed_IvE_5km_sigma50km_ms"], flush=True)
if args.only == "smooth":
    print(json.dumps(res)); sys.exit(0)
rm = gcm.regrid_matrices(sh.name, em, scale=True, correctA=True)
for name, M in (("EvI", rm.matrix("EvI")), ("AvI", rm.matrix("AvI")), ("IvE", rm.matrix("IvE")), ("IvA", rm.matrix("IvA")), ("IvE_smoothed", IvEs)):
    X = torch.rand((M.shape[1], nvar), dtype=torch.float64, device=dev)
    Y = torch.empty((M.shape[0], nvar), dtype=torch.float64, device=dev)
    ms = ev_time(lambda: M.apply_dev(X.data_ptr(), nvar, nvar, Y.data_ptr(), nvar, True, S), n=50, warm=5)
    byt = 12 * M.nnz + 4 * (M.shape[0] + 1) + 8 * nvar * (M.shape[1] + M.shape[0])
    res["apply"][name] = {"shape": list(M.shape), "nnz": M.nnz, "ms": ms, "Gnnz_s": M.nnz * nvar / ms / 1e6, "alg_GBs": byt / ms / 1e6,
                          "frac_of_measured_peak": byt / ms / 1e6 / PEAK, "force_conservation": not M.conservative}
    print("apply", name, res["apply"][name], flush=True)
print(json.dumps(res))
if args.out:
    json.dump(res, open(args.out, "w"), indent=1)
