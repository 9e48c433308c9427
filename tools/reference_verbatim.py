"""Times the UNMODIFIED reference (`/root/reference/scripts/semantic_gating/place_recognition.py`, loaded by file
path through oracle/ref_loader.py) on the benchmark's synthetic inputs, on this machine's host cores.

    python tools/reference_verbatim.py [--out profiles/r02_reference_verbatim.json]

The reference's `find_loop_closures` (place_recognition.py:851-911) builds the N x N fp32 matrix and runs an
O(N^2) Python loop for the temporal mask (:882-885), so it runs in full only at BASELINE config 1 (5k x 512-d); for
config 2's shape (4096-d) it is timed on the first 5 000 keyframes, labelled as a sub-problem.  bench.py runs the same
function when /root/reference exists on the box and otherwise cites the file this script wrote.
Test / measurement infrastructure: nothing under the product package imports it.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "multi-level-indoor-slam_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np


def time_find_loop_closures(n: int, d: int, floors: int, k: int = 25, threshold: float = 0.5, gap: float = 10.0,
                            repeats: int = 1):
    """-> dict with seconds, pairs/s, queries/s and the candidate counts of the reference's own find_loop_closures."""
    from oracle import ref_loader
    from semgate import synthetic
    desc, ts, fl = synthetic.make_case(n, d, floors, seed=0)
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        spr, matches = ref_loader.run_find_loop_closures(desc, ts, fl, threshold, gap, k=k, enable_floor_gating=True)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    valid = sum(1 for m in matches if m.is_valid)
    return {"n": n, "d": d, "floors": floors, "k": k, "seconds": best, "pairs_per_s": float(n) * n / best,
            "queries_per_s": n / best, "matches": len(matches), "valid": valid, "cross_floor": len(matches) - valid}


def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        return max([i.get("num_threads", 0) for i in threadpool_info() if i.get("user_api") == "blas"] or [0])
    except Exception:
        return 0


def measure(sub_n: int = 5000):
    from oracle import ref_loader
    if not ref_loader.available():
        return None
    out = {"what": "unmodified reference SemanticPlaceRecognition.find_loop_closures(enable_floor_gating=True, k=25), "
                   "loaded by path; single Python thread + OpenBLAS sgemm",
           "host_cpus": os.cpu_count(), "blas_threads": blas_threads(),
           "c1_full": time_find_loop_closures(5000, 512, 3),
           "c2_subproblem": dict(time_find_loop_closures(sub_n, 4096, 3),
                                 note=f"first {sub_n} keyframes of config 2's shape (the full 20k x 20k fp32 matrix is 1.6 GB and "
                                      "the Python mask loop 16x longer); pairs/s of this sub-problem, not extrapolated")}
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    r = measure()
    if r is None:
        print(json.dumps({"unavailable": "/root/reference is not present on this machine"}))
        sys.exit(0)
    r["where"] = "build container (no GPU)" if not os.path.exists("/dev/nvidia0") else "GPU box"
    s = json.dumps(r, indent=1)
    print(s)
    if a.out:
        with open(a.out, "w") as f:
            f.write(s + "\n")
