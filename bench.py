#!/usr/bin/env python
"""bench.py -- headline benchmark of the IceBin regridding hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--nvar 512]

Workload (config.workload): BASELINE.json configs[2] on the configs[1] grids -- ModelE 2x2.5 (144x90) A grid +
SeaRISE-shaped Greenland 5 km ice grid (301x561), 40 elevation classes; one STEP pushes a batch of 512 fp64 fields
defined on the ice grid through EvI and AvI (scale=true, correctA=true), i.e. the I->E and I->A half of a coupling
step (IceCoupler.cpp:445-446).  value = (nnz(EvI)+nnz(AvI)) * nvar / time, in Gnnz/s, whole job.
With N>1 (torchrun, one rank per GPU) every rank owns one ice sheet (weak scaling) and the sheets' A-/E-grid
contributions are merged with an NCCL all-reduce (multivec.cpp:35-81); see DESIGN.md.

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


# dram__bytes_read.sum + dram__bytes_write.sum of one EvI spmm_tma_kernel launch at the default workload
# (ncu --set full capture, profiles/r1_spmm_tma_full_raw.csv): 284.70 MB + 22.02 MB
NCU_TRAFFIC_EVI = 306718464


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
    return ap.parse_args()


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def make_config(shape, rank=0):
    """Per-rank sheet: same shape, different seed and anchor so that neighbouring ranks share boundary A cells."""
    from icebin_b200 import synth
    hc = synth.hcdefs_conserv(40)
    im, jm = 144, 90
    ids = np.arange(im * jm, dtype=np.int64)
    gA = synth.GridA(im, jm, im * jm, ids, (synth.A_W * synth.A_H) * (1.0 + 0.08 * np.sin(0.37 * (ids % im) + 0.21 * (ids // im))))
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


def cpu_step(EvI, AvI, Xf):
    """One step of the CPU path on Xf [nfields, nI_dense]; returns seconds."""
    t0 = time.perf_counter()
    EvI.apply(Xf)
    AvI.apply(Xf)
    return time.perf_counter() - t0


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = make_config(args.shape, 0)
    _, EvI, AvI, build_s = cpu_reference_setup(cfg)
    nnz = EvI.nnz + AvI.nnz
    nf = int(min(args.nvar, max(8, args.nvar * 100 // max(1, args.steps))))   # bounded sample per step
    X = np.random.default_rng(0).uniform(-1, 1, (nf, EvI.shape[1]))
    for _ in range(min(args.warmup, 2)):
        cpu_step(EvI, AvI, X)
    t = [cpu_step(EvI, AvI, X) for _ in range(args.steps)]
    tot = sum(t)
    val = nnz * nf * args.steps / tot / 1e9
    sample = "%d of %d fields per step through the oracle's EvI and AvI (Eigen-order CSC x dense), full Greenland 5 km matrices" % (nf, args.nvar)
    line = {"impl": "reference", "metric": "EvI/AvI apply Gnnz/s", "value": val, "unit": "Gnnz/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot / args.steps * (args.nvar / nf),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, EvI.shape, AvI.shape, EvI.nnz, AvI.nnz),
            "cpu_baseline": {"value": val, "unit": "Gnnz/s", "cores": 1, "kind": "port", "sample": sample,
                             "build_s_EvI_AvI": build_s, "host_cores_present": os.cpu_count()},
            "e2e": {"value": val, "unit": "Gnnz/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(args, shE, shA, nnzE, nnzA):
    return {"workload": "configs[2]: %d fp64 fields through EvI and AvI on ModelE 2x2.5 (144x90) + %s (301x561 @5 km), 40 ECs, "
                        "scale+correctA" % (args.nvar, args.shape),
            "nvar": args.nvar, "EvI": {"shape": list(shE), "nnz": nnzE}, "AvI": {"shape": list(shA), "nnz": nnzA},
            "l2": "inputs larger than L2 (X = %.0f MB > 126 MB), no flush" % (shE[1] * args.nvar * 8 / 1e6),
            "parallelism": "1 ice sheet per GPU" if args.gpus > 1 else "single GPU"}


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

    cfg = make_config(args.shape, rank)
    sh = cfg["sheets"][0]
    gcm = api.GCMRegridder(cfg["gridA"], cfg["hcdefs"], True)
    gcm.add_sheet(sh.name, sh)

    # ---- regrid-matrix build (six matrices, scale+correctA), device-resident inputs
    em_host = cfg["elevmask"][0]
    build_ms = []
    for it in range(8):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        rm = gcm.regrid_matrices(sh.name, em_host, scale=True, correctA=True)
        mats = {s: rm.matrix(s) for s in ("EvI", "AvI", "IvE", "IvA", "EvA", "AvE")}
        torch.cuda.synchronize()
        if it >= 3:
            build_ms.append(1e3 * (time.perf_counter() - t0))
    build_med = statistics.median(build_ms)
    if world > 1:           # whole-job build time = slowest rank (every rank builds the matrices of its own sheet)
        tb = torch.tensor([build_med], dtype=torch.float64, device=dev)
        dist.all_reduce(tb, op=dist.ReduceOp.MAX)
        build_med = float(tb.item())
    EvI, AvI = mats["EvI"], mats["AvI"]
    nvar = args.nvar
    nI_d, nE_d, nA_d = EvI.shape[1], EvI.shape[0], AvI.shape[0]
    assert AvI.shape[1] == nI_d
    nnz = EvI.nnz + AvI.nnz

    g = torch.Generator(device=dev); g.manual_seed(1234 + rank)
    X = torch.rand((nI_d, nvar), dtype=torch.float64, device=dev, generator=g) * 2 - 1
    YE = torch.empty((nE_d, nvar), dtype=torch.float64, device=dev)
    YA = torch.empty((nA_d, nvar), dtype=torch.float64, device=dev)

    # multi-sheet merge (N>1): as in the coupler (IceCoupler.cpp:371-377, 445-457) the sheets apply UNSCALED EvI/AvI and
    # the GCM side divides the summed contributions by the summed weights (multivec.cpp:35-81)
    mergeA = None
    if world > 1:
        from icebin_b200 import multi
        rm_u = gcm.regrid_matrices(sh.name, em_host, scale=False, correctA=True)
        EvI, AvI = rm_u.matrix("EvI"), rm_u.matrix("AvI")
        assert (EvI.shape, AvI.shape) == ((nE_d, nI_d), (nA_d, nI_d))
        # two merge contexts + two sets of result buffers: the exchange of step s runs on a side stream while the
        # SpMMs of step s+1 run on the main one (batches of fields are independent)
        merges = [multi.SheetMerge([(gcm.nE, EvI), (gcm.nA, AvI)], nvar, dev) for _ in range(2)]
        mergeA = merges[0]
        side = torch.cuda.Stream(device=dev)
        YEs = [YE, torch.empty_like(YE)]
        YAs = [YA, torch.empty_like(YA)]
        spmm_done = [torch.cuda.Event() for _ in range(2)]
        merge_done = [torch.cuda.Event() for _ in range(2)]

    YE_p, YA_p = YE.data_ptr(), YA.data_ptr()
    # pre-built argument arrays for the C ABI (the per-step Python cost matters at a 110 us step)
    import ctypes as C
    from icebin_b200._lib import vp as _vp
    _W2 = (_vp * 2)(EvI._h, AvI._h)
    _ld2 = (C.c_int64 * 2)(nvar, nvar)
    _Yp = [(_vp * 2)(YE_p, YA_p)]
    if mergeA is not None:
        _Yp = [(_vp * 2)(YEs[k].data_ptr(), YAs[k].data_ptr()) for k in range(2)]
    _apply_many = lib().ib_w_apply_many_dev
    _X, _st = _vp(X.data_ptr()), _vp(stream)
    step_no = [0]
    fresh = [True, True]       # no exchange pending on buffer set k (start of the run / of a captured graph)

    def step():
        # EvI and AvI act on the same ice-grid fields (IceCoupler.cpp:445-446): one batched call, the two independent
        # SpMMs run concurrently on forked streams and join back on `stream`
        if mergeA is None:
            check(_apply_many(2, _W2, _X, nvar, nvar, _Yp[0], _ld2, 1, _st))
            return
        k = step_no[0] & 1
        step_no[0] += 1
        if not fresh[k]:
            tstream.wait_event(merge_done[k])             # buffers k are free again
        fresh[k] = False
        check(_apply_many(2, _W2, _X, nvar, nvar, _Yp[k], _ld2, 1, _st))
        spmm_done[k].record(tstream)
        with torch.cuda.stream(side):
            side.wait_event(spmm_done[k])
            merges[k].run([YEs[k], YAs[k]], side.cuda_stream)
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

    # CUDA graph of GSTEPS consecutive steps (launch-bound otherwise: ~10 launches + a collective per 110 us step).
    # The library's forked streams, its stream-ordered allocations and the NCCL all-reduce are all capturable.
    GSTEPS = 8
    graph, launches_per_graph = None, None
    # single-GPU only: with NCCL inside the capture one rank falling back to eager while another replays would
    # desynchronise the collectives, and the few percent it buys at N>1 are not worth that risk
    if not args.no_graph and args.steps >= GSTEPS and world == 1:
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
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
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
    # ---- roofline of the dominant kernel: the EvI apply alone, bracketed by events, same number of steps
    for i in range(args.steps):
        kev[i][0].record()
        EvI.apply_dev(X.data_ptr(), nvar, nvar, YE_p, nvar, True, stream)
        kev[i][1].record()
    torch.cuda.synchronize()
    sampler.stop_flag = True
    kms = [a.elapsed_time(b) for a, b in kev]
    evi_ms = statistics.mean(kms)

    # ---- the I-side applies of the same matrix set (north star: "batched AvI/EvI/IvE application"), each alone, CUDA
    # events around a loop of applies; reported beside the headline, not part of it
    other = None
    if world == 1 and not args.no_e2e:
        try:
            other = {}
            peak_o, _ = measured_peak()
            YI = torch.empty((nI_d, nvar), dtype=torch.float64, device=dev)
            for name, Xin in (("IvE", YE), ("IvA", YA)):
                M = mats[name]
                assert M.shape == (nI_d, Xin.shape[0])
                for _ in range(3):
                    M.apply_dev(Xin.data_ptr(), nvar, nvar, YI.data_ptr(), nvar, True, stream)
                oe = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                nrep = max(3, min(args.steps, 100))
                oe[0].record()
                for _ in range(nrep):
                    M.apply_dev(Xin.data_ptr(), nvar, nvar, YI.data_ptr(), nvar, True, stream)
                oe[1].record()
                torch.cuda.synchronize()
                oms = oe[0].elapsed_time(oe[1]) / nrep
                oalg = 12 * M.nnz + 4 * (M.shape[0] + 1) + 8 * nvar * (M.shape[0] + M.shape[1])
                other[name] = {"shape": list(M.shape), "nnz": M.nnz, "ms": oms, "Gnnz_s": M.nnz * nvar / oms / 1e6,
                               "alg_GBs": oalg / oms / 1e6, "frac_of_peak": oalg / oms / 1e6 / peak_o}
            del YI
        except Exception as e:      # never let the side measurement break the contract line
            sys.stderr.write("bench.py: I-side applies not measured (%s)\n" % str(e).splitlines()[0])
            other = None

    # ---- e2e: the reference-facing host-buffer call (ib_w_apply), H2D of X and D2H of Y inside the timed region
    e2e = None
    if world == 1 and not args.no_e2e:
        Xh = torch.empty((nI_d, nvar), dtype=torch.float64, pin_memory=True)
        Xh.copy_(X.cpu())
        YEh = torch.empty((nE_d, nvar), dtype=torch.float64, pin_memory=True)
        YAh = torch.empty((nA_d, nvar), dtype=torch.float64, pin_memory=True)
        L = lib()

        import ctypes as C
        Wp = (C.c_void_p * 2)(EvI._h, AvI._h)
        Yp = (C.c_void_p * 2)(YEh.data_ptr(), YAh.data_ptr())
        ldp = (C.c_int64 * 2)(nvar, nvar)

        def e2e_step():
            check(L.ib_w_apply_many(2, Wp, Xh.data_ptr(), nvar, nvar, Yp, ldp, 0, 1))
        for _ in range(2):
            e2e_step()
        ne = max(3, min(args.steps, 20))
        t0 = time.perf_counter()
        for _ in range(ne):
            e2e_step()
        te = (time.perf_counter() - t0) / ne
        e2e = {"value": nnz * nvar / te / 1e9, "unit": "Gnnz/s", "ms_per_step": 1e3 * te,
               "h2d_bytes_per_step": nI_d * nvar * 8, "d2h_bytes_per_step": (nE_d + nA_d) * nvar * 8,
               "api": "ib_w_apply_many([EvI, AvI]) with pinned host buffers: X uploaded once, both results downloaded"}
        assert torch.equal(YEh.to(dev), YE), "e2e result differs from the device-resident result"

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_src = measured_peak()
    alg = 12 * EvI.nnz + 4 * (nE_d + 1) + 8 * nvar * (nI_d + nE_d)
    achieved = alg / (evi_ms * 1e-3) / 1e9
    if mergeA is not None:
        line_extra = {"merge": {"shared_E_cells": int(mergeA.parts[0]["nshared"]), "shared_A_cells": int(mergeA.parts[1]["nshared"]),
                                "collective": "NCCL all-reduce(sum) of the shared rows' unscaled sums on a side stream, overlapping the next step's SpMMs, %d B per step" %
                                              (mergeA.nshared * nvar * 8)}}
    else:
        line_extra = {}
    line = {"metric": "EvI/AvI apply Gnnz/s", "value": nnz_total * nvar * args.steps / (ms * 1e-3) / 1e9, "unit": "Gnnz/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, EvI.shape, AvI.shape, EvI.nnz, AvI.nnz),
            "build_ms": {"six_matrices_median": build_med, "min": min(build_ms), "ranks": world,
                         "what": "regrid_matrices + matrix() x {EvI,AvI,IvE,IvA,EvA,AvE}, scale+correctA, host elevmask upload included"},
            "roofline": {"bound": "hbm", "kernel": "spmm_tma_kernel<pair> + spmm_fixup_list_kernel (the EvI apply)", "achieved": achieved,
                         "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak,
                         "frac_of_nominal_8TBs": achieved / 8000.0, "algorithmic_bytes": alg, "ms_per_launch": evi_ms,
                         "timing": "EvI apply alone, CUDA events around each of `steps` applies issued right after the timed region",
                         "traffic": NCU_TRAFFIC_EVI if (args.shape == "greenland_5km" and nvar == 512) else None,
                         "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of spmm_tma_kernel<1>, profiles/r1_spmm_tma_full_raw.csv"},
            "gpu_launches": int(launches), "issue": ("CUDA graph of %d steps" % GSTEPS) if graph is not None else "eager",
            "clocks": sampler.result()}
    line.update(line_extra)
    if e2e is not None:
        line["e2e"] = e2e
    if other:
        line["other_applies"] = other
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
