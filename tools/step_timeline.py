"""Reads the <prefix>_rank<r>.npz files of `bench.py --trace <prefix>` (IB_STEP_TRACE library) and prints, per rank, the
medians over the timed steps of: apply span (first CTA released .. fix-up end), step period, and where the exchange's two
kernels spend their time (launch after the apply, waiting for the peers' counters, storing, ...).  Times in us."""
import glob, json, sys
import numpy as np
prefix = sys.argv[1]
for f in sorted(glob.glob(prefix + "_rank*.npz")):
    d = np.load(f)
    st = d["stream"].astype(np.float64)
    n = len(st)
    ok = st[:, 1] > 0
    rel, mend, frel, fend = st[:, 0], st[:, 1], st[:, 2], st[:, 3]
    out = {"file": f.split("/")[-1], "steps": int(n), "ms_total": float(d["ms_total"][0]), "us_per_step": round(float(d["ms_total"][0]) * 1e3 / n, 2)}
    q = lambda a: [round(float(np.percentile(a[ok[: len(a)]] if len(a) == n else a, p)) / 1e3, 1) for p in (10, 50, 90, 99)]
    out["apply_main_us"] = q(mend - rel)
    out["apply_with_fixup_us"] = q(fend - rel)
    out["period_us"] = q(np.diff(rel))
    out["gap_fixup_end_to_next_release_us"] = q(rel[1:] - fend[:-1])
    if "xchg" in d:
        x = d["xchg"].astype(np.float64)
        out["publish_start_after_fixup_end_us"] = q(x[:, 0] - fend)
        out["publish_wait_done_us"] = q(x[:, 1] - x[:, 0])
        out["publish_store_us"] = q(x[:, 2] - x[:, 1])
        out["publish_fence_flags_us"] = q(x[:, 3] - x[:, 2])
        out["gather_start_after_publish_us"] = q(x[:, 4] - x[:, 3])
        out["gather_wait_ready_us"] = q(x[:, 5] - x[:, 4])
        out["gather_sum_us"] = q(x[:, 6] - x[:, 5])
        out["fixup_end_to_gather_end_us"] = q(x[:, 6] - fend)
    print(json.dumps(out))
