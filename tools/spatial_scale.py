"""Radius join over poses (semgate_spatial_candidates_host, exact fp64 brute force) at growing pose counts.
    python tools/spatial_scale.py"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multi-level-indoor-slam_b200"))
from semgate import _native  # noqa: E402

eng = _native.get_engine(0)
rng = np.random.default_rng(0)
for n in (19000, 100000, 300000, 1000000):
    # a building-sized random walk: 60 m x 60 m x 5 floors, 0.1 m steps -> revisits everywhere
    step = rng.normal(size=(n, 3)) * np.array([0.1, 0.1, 0.01])
    pos = np.cumsum(step, axis=0)
    pos[:, 0] = np.abs((pos[:, 0] + 30) % 120 - 60)
    pos[:, 1] = np.abs((pos[:, 1] + 30) % 120 - 60)
    pos[:, 2] = np.round(np.abs((pos[:, 2] + 6) % 24 - 12) / 3.0) * 3.0
    eng.spatial_candidates_host(pos[:1000], 2.0, 100)
    t0 = time.perf_counter()
    oi, oj, od = eng.spatial_candidates_host(pos, 2.0, 100)
    dt = time.perf_counter() - t0
    print(json.dumps({"poses": n, "pairs_found": int(len(oi)), "seconds": round(dt, 4), "pair_tests_per_s": round(n * (n - 1) / 2 * 2 / dt, 1)}), flush=True)
    if len(oi) > 4e8:
        break
