"""Config 2 (20k x 4096-d all-pairs) through the one-call entry point (CUDA graph: K2 + K3 + K4, programmatic dependent
launches, no memset nodes) against the three-call path of the bench's timed step."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multi-level-indoor-slam_b200"))

import numpy as np
import torch

from semgate import _native, synthetic

eng = _native.get_engine(0)
n, d = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (20000, 4096)
desc, ts_h, fl_h = synthetic.make_case(n, d, 3, seed=0)
x = eng.normalize_cast(torch.from_numpy(desc).cuda())
ts, fl = torch.from_numpy(ts_h).cuda(), torch.from_numpy(fl_h.astype(np.int32)).cuda()
p = _native.make_params(k=25, similarity_threshold=0.5, min_time_gap=10.0, max_floor_diff=0)
eng.set_option("profile", 1)
side = torch.cuda.Stream()
out = {"n": n, "d": d}
with torch.cuda.stream(side):
    for name, fn in (("one_call_graph", lambda: eng.find_loop_closures_device(x, p, ts=ts, floor=fl, use_graph=True)),
                     ("one_call_eager", lambda: eng.find_loop_closures_device(x, p, ts=ts, floor=fl, use_graph=False)),
                     ("three_calls", lambda: eng.compact(eng.gated_topk(x, x, p, q_ts=ts, db_ts=ts, q_floor=fl, db_floor=fl)))):
        for _ in range(5):
            r = fn()
        side.synchronize()
        eng.profile_read()
        reps = 40
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(side)
        for _ in range(reps):
            r = fn()
        e1.record(side)
        side.synchronize()
        k2_ms, k2_n = eng.profile_read()
        out[name] = {"ms_per_step": round(e0.elapsed_time(e1) / reps, 4), "k2_ms": round(k2_ms / max(k2_n, 1), 4), "k2_launches_timed": k2_n,
                     "candidates": int(r[4].item())}
print(json.dumps(out))
