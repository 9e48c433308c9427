"""K2 in the epilogue-bound regime (short descriptors): one vs two epilogue sets, with and without hits.
    python tools/epi_sets_ab.py"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multi-level-indoor-slam_b200"))
from semgate import _native, synthetic  # noqa: E402

eng = _native.get_engine(0)
eng.set_option("profile", 1)
for n, d in ((5000, 512), (5000, 128), (5000, 1024), (20000, 512)):
    x = synthetic.make_descriptors_device(n, d, "cuda", seed=0)
    xb = eng.normalize_cast(x)
    ts = torch.from_numpy(synthetic.make_timestamps(n)).cuda()
    fl = torch.from_numpy(synthetic.make_floors(n, 3).astype(np.int32)).cuda()
    for thr in (0.5, 2.0):
        row = {"n": n, "d": d, "threshold": thr}
        for sets in ("1", "2"):
            os.environ["SEMGATE_EPI_SETS"] = sets
            p = _native.make_params(k=25, similarity_threshold=thr, min_time_gap=10.0, max_floor_diff=0)
            for _ in range(3):
                r = eng.gated_topk(xb, xb, p, q_ts=ts, db_ts=ts, q_floor=fl, db_floor=fl)
            torch.cuda.synchronize()
            eng.profile_read()
            for _ in range(10):
                r = eng.gated_topk(xb, xb, p, q_ts=ts, db_ts=ts, q_floor=fl, db_floor=fl)
            torch.cuda.synchronize()
            ms, cnt = eng.profile_read()
            row["k2_us_sets" + sets] = round(ms / max(cnt, 1) * 1e3, 2)
            row["cand_sets" + sets] = int(r.count.sum().item())
        print(json.dumps(row), flush=True)
os.environ.pop("SEMGATE_EPI_SETS", None)
