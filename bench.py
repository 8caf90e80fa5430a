#!/usr/bin/env python
"""bench.py -- headline benchmark of the IceBin regridding hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--nvar 512]

Workload (config.workload): BASELINE.json configs[2] on the configs[1] grids -- ModelE 2x2.5 (144x90) A grid +
SeaRISE-shaped Greenland 5 km ice grid (301x561), 40 elevation classes; one STEP pushes a batch of 512 fp64 fields
defined on the ice grid through EvI and AvI (scale=true, correctA=true), i.e. the I->E and I->A half of a coupling
step (IceCoupler.cpp:445-446), as ONE call of ib_w_apply_many_dev.  value = (nnz(EvI)+nnz(AvI)) * nvar / time, in
Gnnz/s, whole job.  With N>1 (torchrun, one rank per GPU) every rank owns one ice sheet (weak scaling) and the
sheets' A-/E-grid contributions are merged across ranks (multivec.cpp:35-81); see DESIGN.md.

The same line also carries the rest of BASELINE.json's metric, measured in the same run (`more`): six-matrix build
times (5 km and Greenland 1 km), the other applies (IvE, IvA, sigma-smoothed IvE), first-apply (plan) costs, a
coupler-shaped step (rebuild + 26 fields), Hntr at ETOPO1 scale; at N>1 the strong-scaling build of configs[4].

Prints ONE JSON line on rank 0.  `--impl reference` times the CPU restatement of the reference (oracle/, single
thread: the reference is serial) on the same workload.
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel at the default workload, from the
# committed `ncu --set full` capture (see profiles/README.md); None until captured for the current kernel
NCU_TRAFFIC = {"kernel": "spmm_stream_kernel<3>", "bytes": None, "source": None}
try:
    _t = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))
    NCU_TRAFFIC = _t
except Exception:
    pass


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nvar", type=int, default=512)
    ap.add_argument("--shape", default="greenland_5km")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="issue every step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--no-e2e", action="store_true", help="kernel experiments only; the driver never passes this")
    ap.add_argument("--no-more", action="store_true", help="skip the secondary measurements (`more`)")
    ap.add_argument("--merge", default="peer", choices=["peer", "nccl"],
                    help="N>1: merge of the sheets' A/E contributions by the library's peer-memory kernels (default) or an NCCL all-reduce")
    ap.add_argument("--trace", default="", help="diagnostics: with the IB_STEP_TRACE library, write <prefix>_rank<r>.npz (per-step time stamps)")
    ap.add_argument("--no-merge", action="store_true", help="diagnostics, N>1: skip the merge (independent replicas)")
    ap.add_argument("--spare", type=int, default=24, help="N>1: CTA slots the apply leaves free for the exchange kernels")
    ap.add_argument("--xchg-blocks", type=int, default=64, help="N>1: 128-thread blocks of the exchange kernels")
    ap.add_argument("--nbuf", type=int, default=8, help="N>1: result buffer sets (how many steps a merge may lag its apply)")
    return ap.parse_args()


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def gridA_global():
    from icebin_b200 import synth
    im, jm = 144, 90
    ids = np.arange(im * jm, dtype=np.int64)
    return synth.GridA(im, jm, im * jm, ids, (synth.A_W * synth.A_H) * (1.0 + 0.08 * np.sin(0.37 * (ids % im) + 0.21 * (ids // im))))


def make_config(shape, rank=0):
    """Per-rank sheet: same shape, different seed and anchor so that neighbouring ranks share boundary A cells."""
    from icebin_b200 import synth
    hc = synth.hcdefs_conserv(40)
    gA = gridA_global()
    im = gA.im
    nx, ny, dx = synth.SHEET_SHAPES[shape]
    npx = int(np.ceil(nx * dx * 1e3 / synth.A_W)) + 1
    shift = max(1, int(0.62 * npx))          # neighbouring sheets' ice overlaps in a few columns of A cells
    anchor = (2 + (rank * shift) % (im - npx - 4), 40)
    _, sh = synth.make_gcm_and_sheet(shape, anchor=anchor, gridA=gA, name="sheet%d" % rank)
    em = synth.make_elevmask(sh, hc[-1], seed=rank)
    return dict(gridA=gA, sheets=[sh], hcdefs=hc, elevmask=[em])


class ClockSampler(threading.Thread):
    """Polls SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.active, self.stop_flag = index, [], set(), False, False
        self.idle = []
        self.sm_max = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown", nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown", nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        while not self.stop_flag:
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                (self.samples if self.active else self.idle).append(mhz)
                if self.active:
                    for bit, nm in names.items():
                        if r & bit:
                            self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.001)

    def result(self):
        src = self.samples if self.samples else self.idle
        return {"sm_mhz": statistics.median(src) if src else None, "sm_max_mhz": self.sm_max,
                "reasons": sorted(self.reasons), "samples": len(self.samples),
                "note": "sampled via NVML every ~1 ms during the timed region" if self.samples else "timed region shorter than one NVML poll; warm-up samples"}


def cpu_reference_setup(cfg):
    from oracle import orc
    gA = cfg["gridA"]
    g = orc.Gcm(gA.nA, gA.dim, gA.native_area, cfg["hcdefs"])
    sh = cfg["sheets"][0]
    g.add_sheet(sh.nI, sh.dimI, sh.native_areaI, sh.centroid_xy, sh.ex_ij, sh.ex_area, sh.gridA_proj_area)
    t0 = time.perf_counter()
    EvI = g.matrix(0, cfg["elevmask"][0], "EvI", scale=True, correctA=True)
    AvI = g.matrix(0, cfg["elevmask"][0], "AvI", scale=True, correctA=True)
    build_s = time.perf_counter() - t0
    return g, EvI, AvI, build_s


def cpu_step(EvI, AvI, Xf, pool=None, nthreads=1):
    """One step of the CPU path on Xf [nfields, nI_dense]; returns seconds.  The reference applies one matrix to one
    batch of fields on one thread (Eigen sparse x dense); fields are independent, so with a pool the batch is cut into
    contiguous slices, one per host thread (the C++ oracle runs outside the GIL)."""
    t0 = time.perf_counter()
    if pool is None or nthreads <= 1:
        EvI.apply(Xf)
        AvI.apply(Xf)
    else:
        cuts = [Xf.shape[0] * k // nthreads for k in range(nthreads + 1)]
        parts = [Xf[a:b] for a, b in zip(cuts[:-1], cuts[1:]) if b > a]
        list(pool.map(lambda P: (EvI.apply(P), AvI.apply(P)), parts))
    return time.perf_counter() - t0


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = make_config(args.shape, 0)
    _, EvI, AvI, build_s = cpu_reference_setup(cfg)
    nnz = EvI.nnz + AvI.nnz
    # all the host threads the path can use: the fields of a batch are independent
    try:
        ncpu = len(os.sched_getaffinity(0))
    except AttributeError:
        ncpu = os.cpu_count() or 1
    nthreads = max(1, min(ncpu, 64))
    # bounded sample per step: at least 8 fields per thread, at most the whole batch
    nf = int(min(args.nvar, max(8 * nthreads, args.nvar * 100 * nthreads // max(1, args.steps))))
    nthreads = min(nthreads, max(1, nf // 8))
    X = np.random.default_rng(0).uniform(-1, 1, (nf, EvI.shape[1]))
    from concurrent.futures import ThreadPoolExecutor
    pool = ThreadPoolExecutor(max_workers=nthreads) if nthreads > 1 else None
    for _ in range(min(args.warmup, 2)):
        cpu_step(EvI, AvI, X, pool, nthreads)
    t = [cpu_step(EvI, AvI, X, pool, nthreads) for _ in range(args.steps)]
    tot = sum(t)
    val = nnz * nf * args.steps / tot / 1e9
    sample = ("%d of %d fields per step through the oracle's EvI and AvI (Eigen-order CSC x dense), full Greenland 5 km matrices, "
              "the fields cut into %d slices on %d host threads" % (nf, args.nvar, nthreads, nthreads))
    line = {"impl": "reference", "metric": "EvI/AvI apply Gnnz/s", "value": val, "unit": "Gnnz/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot / args.steps * (args.nvar / nf),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, EvI.shape, AvI.shape, EvI.nnz, AvI.nnz),
            "cpu_baseline": {"value": val, "unit": "Gnnz/s", "cores": nthreads, "kind": "port", "sample": sample,
                             "build_s_EvI_AvI": build_s, "host_cores_present": os.cpu_count()},
            "e2e": {"value": val, "unit": "Gnnz/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(args, shE, shA, nnzE, nnzA):
    return {"workload": "configs[2]: %d fp64 fields through EvI and AvI on ModelE 2x2.5 (144x90) + %s (301x561 @5 km), 40 ECs, "
                        "scale+correctA" % (args.nvar, args.shape),
            "nvar": args.nvar, "EvI": {"shape": list(shE), "nnz": nnzE}, "AvI": {"shape": list(shA), "nnz": nnzA},
            "l2": "inputs larger than L2 (X = %.0f MB > 126 MB), no flush" % (shE[1] * args.nvar * 8 / 1e6),
            "parallelism": "1 ice sheet per GPU" if args.gpus > 1 else "single GPU"}


# ---------------------------------------------------------------------------------------------------------------
def ev_time(torch, fn, n, warm=3):
    """Average device time of fn() over n back-to-back calls (CUDA events on the current stream), ms."""
    for _ in range(warm):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def timed_build(torch, gcm, name, em, specs, reps, **kw):
    """em: host array (uploaded inside the timed region, as a host caller's would be) or a device tensor (resident)."""
    ts = []
    on_dev = hasattr(em, "data_ptr")
    for it in range(reps + 2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        rm = gcm.regrid_matrices_dev(name, em.data_ptr(), **kw) if on_dev else gcm.regrid_matrices(name, em, **kw)
        mats = {s: rm.matrix(s) for s in specs}
        torch.cuda.synchronize()
        if it >= 2:
            ts.append(1e3 * (time.perf_counter() - t0))
        del mats, rm
    return statistics.median(ts), min(ts)


SIX = ("EvI", "AvI", "IvE", "IvA", "EvA", "AvE")


def more_single_gpu(torch, api, args, cfg, gcm, mats, X, YE, YA, stream, peak, nrep):
    """The rest of BASELINE.json's metric on one GPU.  Every entry is guarded: a failure is recorded, never fatal."""
    from icebin_b200 import synth
    dev = X.device
    nvar = args.nvar
    sh, em = cfg["sheets"][0], cfg["elevmask"][0]
    out = {}

    def guarded(key, fn):
        try:
            out[key] = fn()
        except Exception as e:      # never let a side measurement break the contract line
            out[key] = {"error": str(e).splitlines()[0][:200]}
            sys.stderr.write("bench.py: %s not measured (%s)\n" % (key, out[key]["error"]))

    def apply_stats(M, ms, x_once_rows=None):
        alg = 12 * M.nnz + 4 * (M.shape[0] + 1) + 8 * nvar * (M.shape[0] + M.shape[1])
        return {"shape": list(M.shape), "nnz": M.nnz, "ms": ms, "Gnnz_s": M.nnz * nvar / ms / 1e6, "algorithmic_bytes": alg,
                "alg_GBs": alg / ms / 1e6, "frac_of_peak": alg / ms / 1e6 / peak}

    # ---- the other applies of the matrix set, each alone
    def applies():
        res = {}
        YI = torch.empty((mats["IvE"].shape[0], nvar), dtype=torch.float64, device=dev)
        for name, Xin, Yout in (("EvI", X, YE), ("AvI", X, YA), ("IvE", YE, YI), ("IvA", YA, YI)):
            M = mats[name]
            ms = ev_time(torch, lambda: M.apply_dev(Xin.data_ptr(), nvar, nvar, Yout.data_ptr(), nvar, True, stream), nrep)
            res[name] = apply_stats(M, ms)
        # configs[2] "with sigma smoothing enabled": the smoothed IvE, conservation correction included
        rms = gcm.regrid_matrices(sh.name, em, scale=True, correctA=True, sigma=(50e3, 50e3, 100.0))
        t0 = time.perf_counter()
        IvEs = rms.matrix("IvE")
        torch.cuda.synchronize()
        bms = 1e3 * (time.perf_counter() - t0)
        ms = ev_time(torch, lambda: IvEs.apply_dev(YE.data_ptr(), nvar, nvar, YI.data_ptr(), nvar, True, stream), max(3, nrep // 4))
        res["IvE_smoothed"] = dict(apply_stats(IvEs, ms), sigma=[50e3, 50e3, 100.0], build_ms=bms, force_conservation=True)
        ms = ev_time(torch, lambda: IvEs.apply_dev(YE.data_ptr(), nvar, nvar, YI.data_ptr(), nvar, False, stream), max(3, nrep // 4))
        res["IvE_smoothed_no_correction"] = dict(apply_stats(IvEs, ms), force_conservation=False,
                                                 kernel="smooth_group_kernel (DMMA.8x8x4 blocks)")
        return res
    guarded("applies_alone", applies)

    # ---- first-apply cost (plan construction) per matrix: first call minus a steady-state call, fresh matrices
    def plans():
        rm = gcm.regrid_matrices(sh.name, em, scale=True, correctA=True)
        res = {}
        YI = torch.empty((mats["IvE"].shape[0], nvar), dtype=torch.float64, device=dev)
        for name, Xin, Yout in (("EvI+AvI", X, None), ("IvE", YE, YI), ("IvA", YA, YI)):
            if name == "EvI+AvI":
                Ms = [rm.matrix("EvI"), rm.matrix("AvI")]
                fn = lambda: api.apply_many_dev(Ms, X.data_ptr(), nvar, nvar, [YE.data_ptr(), YA.data_ptr()], [nvar, nvar], True, stream)
            else:
                M = rm.matrix(name)
                fn = lambda: M.apply_dev(Xin.data_ptr(), nvar, nvar, Yout.data_ptr(), nvar, True, stream)
            torch.cuda.synchronize()
            t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); t1 = time.perf_counter()
            fn(); torch.cuda.synchronize(); t2 = time.perf_counter()
            res[name] = {"first_apply_ms": 1e3 * (t1 - t0), "next_apply_ms": 1e3 * (t2 - t1), "plan_ms": 1e3 * ((t1 - t0) - (t2 - t1))}
        return res
    guarded("plan_ms", plans)

    # ---- a coupler-shaped step (IceCoupler.cpp:361-468): rebuild E1vI / A1vI (unscaled, correctA) and the smoothless IvE,
    # apply EvI and AvI once to 26 fields and IvE once to 5 -- matrices used once, so plan cost is inside
    def coupler():
        nf_out, nf_in = 26, 5
        Xc = torch.rand((X.shape[0], nf_out), dtype=torch.float64, device=dev)
        Yc = [torch.empty((YE.shape[0], nf_out), dtype=torch.float64, device=dev), torch.empty((YA.shape[0], nf_out), dtype=torch.float64, device=dev)]
        Zc = torch.rand((YE.shape[0], nf_in), dtype=torch.float64, device=dev)
        Ic = torch.empty((X.shape[0], nf_in), dtype=torch.float64, device=dev)
        ts = []
        for it in range(7):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            rm = gcm.regrid_matrices(sh.name, em, scale=False, correctA=True)
            E, A = rm.matrix("EvI"), rm.matrix("AvI")
            I = gcm.regrid_matrices(sh.name, em, scale=True, correctA=True).matrix("IvE")
            api.apply_many_dev([E, A], Xc.data_ptr(), nf_out, nf_out, [y.data_ptr() for y in Yc], [nf_out, nf_out], True, stream)
            I.apply_dev(Zc.data_ptr(), nf_in, nf_in, Ic.data_ptr(), nf_in, True, stream)
            torch.cuda.synchronize()
            if it >= 2:
                ts.append(1e3 * (time.perf_counter() - t0))
        return {"ms_median": statistics.median(ts), "min": min(ts),
                "what": "regrid_matrices x2 + matrix(EvI, AvI unscaled; IvE) + apply_many([EvI, AvI]) on 26 fields + IvE on 5 fields, wall clock, host elevmask upload included"}
    guarded("coupler_shaped_step", coupler)

    # ---- north-star build target: Greenland 1 km x 40 ECs, six matrices
    def build_1km():
        gA = gridA_global()
        _, gr = synth.make_gcm_and_sheet("greenland_1km", anchor=(40, 60), gridA=gA, name="greenland_1km")
        em1 = synth.make_elevmask(gr, cfg["hcdefs"][-1], 0)
        g1 = api.GCMRegridder(gA, cfg["hcdefs"], True)
        g1.add_sheet(gr.name, gr)
        med, mn = timed_build(torch, g1, gr.name, em1, SIX, 5, scale=True, correctA=True)
        em1d = torch.from_numpy(em1).to(dev)
        medd, mnd = timed_build(torch, g1, gr.name, em1d, SIX, 5, scale=True, correctA=True)
        return {"six_matrices_median_ms": med, "min_ms": mn, "nI": gr.nI, "nX": gr.nX, "target_ms": 100.0,
                "what": "regrid_matrices + matrix() x {EvI,AvI,IvE,IvA,EvA,AvE}, scale+correctA, wall clock, host elevmask upload (34 MB, pageable) included",
                "elevmask_resident": {"six_matrices_median_ms": medd, "min_ms": mnd, "what": "the same through ib_regrid_matrices_dev: elevmaskI already in HBM"}}
    guarded("build_greenland_1km", build_1km)

    # ---- Hntr at ETOPO1 scale (make_topoo / etopo1_ice.cpp:17-29): 21600x10800 -> 144x90 and 360x180
    def hntr():
        res = {}
        G = api.GRIDS
        A = G["g1mx1m"]
        src16 = torch.randint(-500, 4000, (A.size,), dtype=torch.int16, device=dev)
        for bname in ("g2hx2", "g1x1"):
            h = api.Hntr(17.17, G[bname], A, 0.0)
            o = torch.empty(G[bname].size, dtype=torch.float64, device=dev)
            ms = ev_time(torch, lambda: h.regrid_dev(0, src16.data_ptr(), 1, o.data_ptr(), True, 1.0, 0.0, stream), 20)
            byt = 2 * A.size + 8 * G[bname].size
            res["g1mx1m->%s int16, constant weight" % bname] = {"ms": ms, "algorithmic_bytes": byt, "alg_GBs": byt / ms / 1e6, "frac_of_peak": byt / ms / 1e6 / peak}
        del src16
        src64 = torch.rand(A.size, dtype=torch.float64, device=dev)
        w64 = (torch.rand(A.size, device=dev) < 0.3).to(torch.float64)
        for bname in ("g2hx2", "g1x1"):
            h = api.Hntr(17.17, G[bname], A, 0.0)
            o = torch.empty(G[bname].size, dtype=torch.float64, device=dev)
            ms = ev_time(torch, lambda: h.regrid_dev(w64.data_ptr(), src64.data_ptr(), 0, o.data_ptr(), True, 1.0, 0.0, stream), 10)
            byt = 16 * A.size + 8 * G[bname].size
            res["g1mx1m->%s fp64, fp64 weights" % bname] = {"ms": ms, "algorithmic_bytes": byt, "alg_GBs": byt / ms / 1e6, "frac_of_peak": byt / ms / 1e6 / peak}
        return res
    guarded("hntr_etopo1", hntr)
    return out


def more_config5_n1(torch, api, args, dev):
    """The N=1 point of the configs[4] strong-scaling build curve (same code path as N>1)."""
    try:
        return strong_scaling_build(torch, None, api, args, 1, 0, dev)
    except Exception as e:
        return {"error": str(e).splitlines()[0][:200]}


def strong_scaling_build(torch, dist, api, args, world, rank, dev):
    """configs[4] at N GPUs: the GCM cells of each 1 km sheet are cut into N domains of equal exchange-cell count
    (multi.partition_cellsA -> GCMRegridder.filter_cellsA, the reference's chunking: modele/global_ec.cpp:818-893); every
    rank builds the six matrices of its domain; the job's time is the slowest rank's median."""
    from icebin_b200 import multi, synth
    hc = synth.hcdefs_conserv(40)
    gA = gridA_global()
    out = {}
    shapes = [("greenland_1km", (40, 60), 5)]
    try:
        import psutil
        per_rank = psutil.virtual_memory().available / max(1, world)
    except Exception:
        per_rank = 0
    if per_rank > 24e9:
        shapes.append(("antarctica_1km", (40, 2), 3))
    else:
        out["antarctica_1km"] = {"skipped": "less than 24 GB of host memory per rank for the synthetic 36 M-cell grid"}
    for shape, anchor, reps in shapes:
        _, sh = synth.make_gcm_and_sheet(shape, anchor=anchor, gridA=gA, name=shape)
        em = synth.make_elevmask(sh, hc[-1], 0)
        g = api.GCMRegridder(gA, hc, True)
        g.add_sheet(sh.name, sh)
        nX_full = g.nX(sh.name)
        if world > 1:
            g.filter_cellsA(multi.partition_cellsA(sh.ex_ij[:, 0], gA.nA, world)[rank])
        nX_mine = g.nX(sh.name)
        del sh
        em = torch.from_numpy(em).to(dev)       # the ice model's elevation / mask lives in HBM: no 288 MB upload per build
        if world > 1:
            dist.barrier()
        med, mn = timed_build(torch, g, shape, em, SIX, reps, scale=False, correctA=True)
        t = torch.tensor([med, float(nX_mine)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out[shape] = {"six_matrices_ms_slowest_rank_median": float(t[0].item()), "nX": nX_full, "nX_largest_domain": int(t[1].item()), "ranks": world}
        del g, em
    out["what"] = ("strong scaling: every rank builds {EvI,AvI,IvE,IvA,EvA,AvE} (unscaled, correctA) for its GCM-cell domain of the sheet, elevmaskI "
                   "resident in HBM (ib_regrid_matrices_dev); rows of AvI/EvI of the domains are disjoint and concatenate (ib_w_combine_chunks)")
    return out


def run_b200(args):
    import torch
    import torch.distributed as dist
    from icebin_b200 import api

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if api.device_count() == 0 or not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; icebin_b200 has no CPU path (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    from icebin_b200._lib import lib, check
    check(lib().ib_set_device(local))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    # a real (non-NULL) stream: the library launches on the stream it is handed, and torch.cuda.Event only sees
    # torch's current stream, so the whole timed region runs under `with torch.cuda.stream(tstream)`
    tstream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    assert stream != 0
    peak, peak_src = measured_peak()

    cfg = make_config(args.shape, rank)
    sh = cfg["sheets"][0]
    gcm = api.GCMRegridder(cfg["gridA"], cfg["hcdefs"], True)
    gcm.add_sheet(sh.name, sh)

    # ---- regrid-matrix build (six matrices, scale+correctA) on the headline grids
    em_host = cfg["elevmask"][0]
    build_med, build_min = timed_build(torch, gcm, sh.name, em_host, SIX, 5, scale=True, correctA=True)
    if world > 1:           # whole-job build time = slowest rank (every rank builds the matrices of its own sheet)
        tb = torch.tensor([build_med], dtype=torch.float64, device=dev)
        dist.all_reduce(tb, op=dist.ReduceOp.MAX)
        build_med = float(tb.item())
    rm = gcm.regrid_matrices(sh.name, em_host, scale=True, correctA=True)
    mats = {s: rm.matrix(s) for s in SIX}
    EvI, AvI = mats["EvI"], mats["AvI"]
    nvar = args.nvar
    nI_d, nE_d, nA_d = EvI.shape[1], EvI.shape[0], AvI.shape[0]
    assert AvI.shape[1] == nI_d
    nnz = EvI.nnz + AvI.nnz

    g = torch.Generator(device=dev); g.manual_seed(1234 + rank)
    X = torch.rand((nI_d, nvar), dtype=torch.float64, device=dev, generator=g) * 2 - 1
    YE = torch.empty((nE_d, nvar), dtype=torch.float64, device=dev)
    YA = torch.empty((nA_d, nvar), dtype=torch.float64, device=dev)

    # multi-sheet merge (N>1): the GCM side sums the sheets' contributions and divides by the summed weights (multivec.cpp:35-81)
    mergeA = None
    if world > 1:
        from icebin_b200 import multi
        # two sets of result buffers: the exchange of step s runs on a side stream while the SpMM of step s+1 runs on the
        # main one (batches of fields are independent)
        if args.merge == "peer":
            # the library's own publish / gather kernels over NVLink peer memory (csrc/xchg.cu), no collective call.  The
            # sheets apply their SCALED matrices, so the rows of unshared cells are final and only the shared cells travel
            # (as sum = Y * wM) and are rewritten: sum over owners / summed weights (multivec.cpp:35-81)
            peer = multi.PeerMerge([(gcm.nE, EvI), (gcm.nA, AvI)], nvar, dev, scaled=True)
            mergeA = peer
            # the small merge kernels need free CTA slots next to the SpMM's resident grid, and no more blocks than slots:
            # a block that finds none waits for the SpMM's CTAs to exit (measured: publish took one whole SpMM, 45 us)
            api.tune("stream_spare", args.spare)
            api.tune("xchg_blocks", args.xchg_blocks)
        else:
            # as in the coupler (IceCoupler.cpp:371-377, 445-457): UNSCALED EvI / AvI, pack + NCCL all-reduce + unpack
            rm_u = gcm.regrid_matrices(sh.name, em_host, scale=False, correctA=True)
            EvI, AvI = rm_u.matrix("EvI"), rm_u.matrix("AvI")
            assert (EvI.shape, AvI.shape) == ((nE_d, nI_d), (nA_d, nI_d))
            merges = [multi.SheetMerge([(gcm.nE, EvI), (gcm.nA, AvI)], nvar, dev) for _ in range(4)]
            mergeA = merges[0]
        side = torch.cuda.Stream(device=dev, priority=-1)       # the small merge kernels go first when a CTA slot frees up
        # publish and gather on streams of their own: next to a running SpMM a publish takes about as long as the SpMM
        # (its loads and its system-scope fence queue behind the SpMM's traffic), so publish(s + 1) has to overlap gather(s)
        side2 = torch.cuda.Stream(device=dev, priority=-1)
        NBUF = max(2, min(8, args.nbuf))            # result buffer sets: a merge may lag its SpMM by NBUF - 1 steps (absorbs the ranks' jitter)
        YEs = [YE] + [torch.empty_like(YE) for _ in range(NBUF - 1)]
        YAs = [YA] + [torch.empty_like(YA) for _ in range(NBUF - 1)]
        spmm_done = [torch.cuda.Event() for _ in range(NBUF)]
        merge_done = [torch.cuda.Event() for _ in range(NBUF)]
        pub_done = [torch.cuda.Event() for _ in range(NBUF)]

    YE_p, YA_p = YE.data_ptr(), YA.data_ptr()
    # pre-built argument arrays for the C ABI (the per-step Python cost matters at a 60 us step)
    import ctypes as C
    from icebin_b200._lib import vp as _vp
    _W2 = (_vp * 2)(EvI._h, AvI._h)
    _ld2 = (C.c_int64 * 2)(nvar, nvar)
    _Yp = [(_vp * 2)(YE_p, YA_p)]
    if mergeA is not None:
        _Yp = [(_vp * 2)(YEs[k].data_ptr(), YAs[k].data_ptr()) for k in range(NBUF)]
    _apply_many = lib().ib_w_apply_many_dev
    _X, _st = _vp(X.data_ptr()), _vp(stream)
    if mergeA is not None and args.merge == "peer":
        _pa = [peer.args([YEs[k], YAs[k]]) for k in range(NBUF)]
        _side = _vp(side.cuda_stream)
        _side2 = _vp(side2.cuda_stream)
    step_no = [0]
    fresh = [True] * 8         # no exchange pending on buffer set k (start of the run / of a captured graph)

    def step():
        # EvI and AvI act on the same ice-grid fields (IceCoupler.cpp:445-446): one call; the library applies the
        # stacked [EvI; AvI] in one pass over X (spmm_stream_kernel<3> + its small fix-up kernel)
        if mergeA is None:
            check(_apply_many(2, _W2, _X, nvar, nvar, _Yp[0], _ld2, 1, _st))
            return
        k = step_no[0] % NBUF
        step_no[0] += 1
        if not fresh[k]:
            tstream.wait_event(merge_done[k])             # buffers k are free again
        fresh[k] = False
        check(_apply_many(2, _W2, _X, nvar, nvar, _Yp[k], _ld2, 1, _st))
        spmm_done[k].record(tstream)
        if args.no_merge:
            merge_done[k].record(tstream)
            return
        side.wait_event(spmm_done[k])
        if args.merge == "peer":
            peer.publish_raw(_pa[k][0], _pa[k][1], nvar, _side)
            pub_done[k].record(side)
            side2.wait_event(pub_done[k])
            peer.gather_raw(_pa[k][0], _pa[k][1], nvar, _side2)
            merge_done[k].record(side2)
        else:
            merges[k].run([YEs[k], YAs[k]], side)
            merge_done[k].record(side)

    def drain():
        if mergeA is not None:
            for k, e in enumerate(merge_done):
                if not fresh[k]:
                    tstream.wait_event(e)
                fresh[k] = True
            step_no[0] = 0

    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        step()
    drain()
    torch.cuda.synchronize()

    # CUDA graph of GSTEPS consecutive steps (launch-bound otherwise).  The library's launches (programmatic dependent
    # launches included) and its stream-ordered allocations are capturable.
    # steps per captured graph: the largest divisor of `steps` in 8..64 (no eager remainder), else 16.  At N>1 a graph ends
    # with the join of the last merges, so longer graphs expose that latency less often.
    divs = [d for d in range(8, 65) if args.steps % d == 0]
    GSTEPS = (max(divs) if world > 1 else min(divs)) if divs else 16
    graph, launches_per_graph = None, None
    # N>1 with the peer merge: nothing but the library's kernels is captured (no collective call), every rank captures the
    # same GSTEPS steps.  With NCCL inside the capture one rank falling back to eager while another replays would
    # desynchronise the collectives, so that arm stays eager.
    if not args.no_graph and args.steps >= GSTEPS and (world == 1 or args.merge == "peer"):
        try:
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            l0 = api.launch_count()
            with torch.cuda.graph(graph, stream=tstream):
                for _ in range(GSTEPS):
                    step()
                drain()
            launches_per_graph = api.launch_count() - l0
            for _ in range(2):
                graph.replay()
            torch.cuda.synchronize()
        except Exception as e:      # capture not supported in this setup: fall back to eager issue
            sys.stderr.write("bench.py: CUDA graph capture failed (%s); running eagerly\n" % str(e).splitlines()[0])
            graph = None
            torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    trace_on = bool(args.trace) and hasattr(lib(), "ib_debug_stream_steps")
    if args.trace and not trace_on:
        sys.stderr.write("bench.py: --trace needs the IB_STEP_TRACE library (tools/build_variants.py steptrace IB_STEP_TRACE)\n")
    if trace_on:
        torch.cuda.synchronize()
        lib().ib_debug_stream_steps.argtypes = [C.c_void_p, C.c_int, C.c_int]
        lib().ib_debug_stream_steps(None, 0, 1)
        x_step0 = peer.steps_done() if (world > 1 and args.merge == "peer") else 0
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    l0 = api.launch_count()
    sampler.active = True
    ev[0].record()
    if graph is not None:
        nrep, rest = divmod(args.steps, GSTEPS)
        for i in range(nrep):
            graph.replay()
        for i in range(rest):
            step()
    else:
        for i in range(args.steps):
            step()
    drain()                     # the last exchanges are inside the timed region
    ev[1].record()
    torch.cuda.synchronize()
    sampler.active = False
    if trace_on:
        import numpy as _np
        n = min(4096, args.steps)
        st = _np.zeros((4096, 4), _np.uint64)
        lib().ib_debug_stream_steps(st.ctypes.data_as(C.c_void_p), 4096, 0)
        out = {"stream": st[:n], "ms_total": _np.array([ev[0].elapsed_time(ev[1])])}
        if world > 1 and args.merge == "peer":
            xs = _np.zeros((4096, 8), _np.uint64)
            lib().ib_debug_xchg_steps.argtypes = [C.c_void_p, C.c_int]
            lib().ib_debug_xchg_steps(xs.ctypes.data_as(C.c_void_p), 4096)
            idx = (_np.arange(n) + x_step0) % 4096
            out["xchg"] = xs[idx]
        _np.savez("%s_rank%d.npz" % (args.trace, rank), **out)
    launches = api.launch_count() - l0
    if graph is not None:
        launches += (args.steps // GSTEPS) * launches_per_graph      # kernels replayed from the graph
    if world > 1:
        dist.barrier()
    ms = ev[0].elapsed_time(ev[1])
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        tn = torch.tensor([float(nnz)], dtype=torch.float64, device=dev)
        dist.all_reduce(tn, op=dist.ReduceOp.SUM)
        nnz_total = float(tn.item())
    else:
        nnz_total = float(nnz)
    # ---- roofline of the dominant kernel (spmm_stream_kernel<3>, the stacked apply), alone: an event pair around every
    # apply, same number of steps, right after the timed region (inside the N>1 step it overlaps the exchange)
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for i in range(args.steps):
        kev[i][0].record()
        check(_apply_many(2, _W2, _X, nvar, nvar, _Yp[0], _ld2, 1, _st))
        kev[i][1].record()
    torch.cuda.synchronize()
    sampler.stop_flag = True
    stack_ms = statistics.mean(a.elapsed_time(b) for a, b in kev)

    nrep_side = max(3, min(args.steps, 100))
    more = None
    if world == 1 and not args.no_more:
        more = more_single_gpu(torch, api, args, cfg, gcm, mats, X, YE, YA, stream, peak, nrep_side)
        more["build_config5_strong_scaling"] = more_config5_n1(torch, api, args, dev)
    if world > 1 and not args.no_more:
        try:
            more = {"build_config5_strong_scaling": strong_scaling_build(torch, dist, api, args, world, rank, dev)}
        except Exception as e:
            more = {"build_config5_strong_scaling": {"error": str(e).splitlines()[0][:200]}}

    # ---- e2e: the reference-facing host-buffer call, H2D of X and D2H of both results inside the timed region
    e2e = None
    if not args.no_e2e:
        Xh = torch.empty((nI_d, nvar), dtype=torch.float64, pin_memory=True)
        Xh.copy_(X.cpu())
        YEh = torch.empty((nE_d, nvar), dtype=torch.float64, pin_memory=True)
        YAh = torch.empty((nA_d, nvar), dtype=torch.float64, pin_memory=True)
        L = lib()
        Wp = (C.c_void_p * 2)(EvI._h, AvI._h)
        ne = max(3, min(args.steps, 20))
        if world == 1:
            def timed(fn):
                for _ in range(2):
                    fn()
                t0 = time.perf_counter()
                for _ in range(ne):
                    fn()
                return (time.perf_counter() - t0) / ne
            Yp = (C.c_void_p * 2)(YEh.data_ptr(), YAh.data_ptr())
            ldp = (C.c_int64 * 2)(nvar, nvar)
            te = timed(lambda: check(L.ib_w_apply_many(2, Wp, Xh.data_ptr(), nvar, nvar, Yp, ldp, 0, 1)))
            assert torch.allclose(YEh.to(dev), YE, rtol=1e-11, atol=1e-12), "e2e result differs from the device-resident result"
            # the reference's own layout, [nvar x ncell] (IceCoupler.cpp:231-233): one device transposition each way
            Xf = torch.empty((nvar, nI_d), dtype=torch.float64, pin_memory=True)
            Xf.copy_(X.t().cpu())
            YEf = torch.empty((nvar, nE_d), dtype=torch.float64, pin_memory=True)
            YAf = torch.empty((nvar, nA_d), dtype=torch.float64, pin_memory=True)
            Ypf = (C.c_void_p * 2)(YEf.data_ptr(), YAf.data_ptr())
            ldf = (C.c_int64 * 2)(nE_d, nA_d)
            tf = timed(lambda: check(L.ib_w_apply_many(2, Wp, Xf.data_ptr(), nI_d, nvar, Ypf, ldf, 1, 1)))
            assert torch.equal(YEf.t().contiguous(), YEh), "field-major e2e result differs from the cell-major one"
            e2e = {"value": nnz * nvar / te / 1e9, "unit": "Gnnz/s", "ms_per_step": 1e3 * te,
                   "h2d_bytes_per_step": nI_d * nvar * 8, "d2h_bytes_per_step": (nE_d + nA_d) * nvar * 8,
                   "api": "ib_w_apply_many([EvI, AvI]) with pinned host buffers, cell-major fields: X uploaded once, both results downloaded",
                   "field_major": {"value": nnz * nvar / tf / 1e9, "ms_per_step": 1e3 * tf,
                                   "note": "the reference's [nvar x ncell] layout (IB_LAYOUT_FIELD_MAJOR): transposed on the device both ways"}}
        else:
            # every rank: upload its sheet's fields, apply, merge across ranks, download -- wall clock, barrier on both sides
            def e2e_step():
                X.copy_(Xh, non_blocking=True)
                check(_apply_many(2, _W2, _X, nvar, nvar, _Yp[0], _ld2, 1, _st))
                if args.merge == "peer":
                    peer.merge_raw(_pa[0][0], _pa[0][1], nvar, _st)
                else:
                    merges[0].run([YEs[0], YAs[0]], tstream)
                YEh.copy_(YEs[0], non_blocking=True)
                YAh.copy_(YAs[0], non_blocking=True)
                tstream.synchronize()
            for _ in range(2):
                e2e_step()
            dist.barrier()
            t0 = time.perf_counter()
            for _ in range(ne):
                e2e_step()
            dist.barrier()
            te = (time.perf_counter() - t0) / ne
            tt = torch.tensor([te], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            te = float(tt.item())
            e2e = {"value": nnz_total * nvar / te / 1e9, "unit": "Gnnz/s", "ms_per_step": 1e3 * te,
                   "h2d_bytes_per_step": nI_d * nvar * 8 * world, "d2h_bytes_per_step": (nE_d + nA_d) * nvar * 8 * world,
                   "api": "per rank: pinned X -> device, ib_w_apply_many_dev([EvI, AvI]), merge across ranks (%s), both results -> pinned host; max over ranks" %
                          ("ib_xchg_merge_dev" if args.merge == "peer" else "SheetMerge.run: NCCL")}

    if world > 1:
        torch.cuda.synchronize()
        dist.barrier()          # nobody unmaps its exchange buffer while a peer may still read it
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    alg = 12 * nnz + 8 * nvar * (nI_d + nE_d + nA_d)
    achieved = alg / (stack_ms * 1e-3) / 1e9
    if mergeA is not None:
        how = ("ib_xchg_publish_dev / ib_xchg_gather_dev: every rank stores the rows of its shared cells into its peers' inboxes over NVLink and adds "
               "up its own inbox (no collective call); publish and gather on two side streams (publish(s + 1) overlaps gather(s)), both overlapping "
               "later steps' SpMMs; the SpMM leaves " + str(args.spare) + " CTA slots free, the exchange kernels run " + str(args.xchg_blocks) +
               " blocks of 128 threads, " + str(NBUF) + " result buffer sets; at most %d B published per rank and step"
               if args.merge == "peer" else
               "NCCL all-reduce(sum) of the shared rows' unscaled sums on a side stream, overlapping the next step's SpMM, %d B per step")
        line_extra = {"merge": {"shared_E_cells": int(mergeA.parts[0]["nshared"]), "shared_A_cells": int(mergeA.parts[1]["nshared"]),
                                "how": how % (mergeA.nshared * nvar * 8)}}
    else:
        line_extra = {}
    default_wl = args.shape == "greenland_5km" and nvar == 512
    line = {"metric": "EvI/AvI apply Gnnz/s", "value": nnz_total * nvar * args.steps / (ms * 1e-3) / 1e9, "unit": "Gnnz/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, EvI.shape, AvI.shape, EvI.nnz, AvI.nnz),
            "build_ms": {"six_matrices_median": build_med, "min": build_min, "ranks": world, "grid": args.shape,
                         "what": "regrid_matrices + matrix() x {EvI,AvI,IvE,IvA,EvA,AvE}, scale+correctA, host elevmask upload included"},
            "roofline": {"bound": "hbm", "kernel": "spmm_stream_kernel<3> (+ spmm_stream_fixup_kernel): the stacked [EvI; AvI] apply, X read once",
                         "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak,
                         "frac_of_nominal_8TBs": achieved / 8000.0, "algorithmic_bytes": alg,
                         "algorithmic_bytes_formula": "12*(nnz_EvI+nnz_AvI) + 8*nvar*(ncol + nrow_EvI + nrow_AvI)", "ms_per_launch": stack_ms,
                         "timing": "one stacked apply per event pair, `steps` of them issued right after the timed region",
                         "traffic": NCU_TRAFFIC.get("bytes") if default_wl else None, "traffic_source": NCU_TRAFFIC.get("source")},
            "gpu_launches": int(launches), "issue": ("CUDA graph of %d steps" % GSTEPS) if graph is not None else "eager",
            "clocks": sampler.result()}
    line.update(line_extra)
    if e2e is not None:
        line["e2e"] = e2e
    if more:
        line["more"] = more
    if world == 1 and not args.no_cpu_baseline:
        _, oE, oA, bs = cpu_reference_setup(cfg)
        nf = 64
        Xf = np.ascontiguousarray(X[:, :nf].T.cpu().numpy())
        cpu_step(oE, oA, Xf)
        reps = 3
        tc = min(cpu_step(oE, oA, Xf) for _ in range(reps))
        line["cpu_baseline"] = {"value": nnz * nf / tc / 1e9, "unit": "Gnnz/s", "cores": 1, "kind": "port",
                                "sample": "%d of %d fields, full EvI+AvI matrices, best of %d; oracle build of EvI+AvI took %.2f s" % (nf, nvar, reps, bs),
                                "host_cores_present": os.cpu_count()}
        # cross-check while we are here: GPU result == oracle result on the sampled fields
        Yo = oE.apply(Xf)
        Yg = YE[:, :nf].T.cpu().numpy()
        line["cpu_baseline"]["max_abs_diff_vs_gpu"] = float(np.abs(Yo - Yg).max())
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
