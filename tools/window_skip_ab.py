"""K2 on a sequence whose temporal neighbours look alike (slow random walk of the descriptor, 10 Hz keyframes, 10 s window):
the chunk-level window test on / off.     python tools/window_skip_ab.py"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multi-level-indoor-slam_b200"))
from semgate import _native  # noqa: E402

eng = _native.get_engine(0)
eng.set_option("profile", 1)
for n, d in ((5000, 512), (20000, 512), (20000, 4096)):
    g = torch.Generator(device="cuda").manual_seed(1)
    places = 40
    # a walk that visits `places` places in turn and comes back: neighbours in time are near copies, revisits are loop closures
    anchors = torch.randn((places, d), device="cuda", generator=g)
    seg = n // (2 * places)
    pid = (torch.arange(n, device="cuda") // seg) % places
    x = anchors[pid] + 0.02 * torch.cumsum(torch.randn((n, d), device="cuda", generator=g), dim=0) % 1.0 + 0.3 * torch.randn((n, d), device="cuda", generator=g)
    xb = eng.normalize_cast(x)
    ts = torch.arange(n, device="cuda", dtype=torch.float64) * 0.1 + 1000.0
    fl = (pid % 3).to(torch.int32)
    row = {"n": n, "d": d}
    for skip in ("1", "0"):
        os.environ["SEMGATE_WINDOW_SKIP"] = skip
        p = _native.make_params(k=25, similarity_threshold=0.5, min_time_gap=10.0, max_floor_diff=0)
        for _ in range(3):
            r = eng.gated_topk(xb, xb, p, q_ts=ts, db_ts=ts, q_floor=fl, db_floor=fl)
        torch.cuda.synchronize()
        eng.profile_read()
        for _ in range(10):
            r = eng.gated_topk(xb, xb, p, q_ts=ts, db_ts=ts, q_floor=fl, db_floor=fl)
        torch.cuda.synchronize()
        ms, cnt = eng.profile_read()
        row["k2_us_skip" + skip] = round(ms / max(cnt, 1) * 1e3, 2)
        row["cand_skip" + skip] = int(r.count.sum().item())
    print(json.dumps(row), flush=True)
os.environ.pop("SEMGATE_WINDOW_SKIP", None)
