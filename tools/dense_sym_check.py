"""Symmetric vs full-matrix sweep on a sequence-like input with dense hits (every place visited twice for `seg` keyframes:
~2 seg hits per row above the threshold): time, whether the symmetric sweep's candidate buffers overflowed (mode 0 = the
armed full sweep redid the job), identical lists.     python tools/dense_sym_check.py"""
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


def seq(n, d, seg, noise=0.3):
    g = torch.Generator(device="cuda").manual_seed(1)
    places = max(2, n // (2 * seg))
    anchors = torch.randn((places, d), device="cuda", generator=g)
    pid = (torch.arange(n, device="cuda") // seg) % places
    out = torch.empty((n, _native.pad_dim(d)), dtype=torch.bfloat16, device="cuda")
    for s0 in range(0, n, 65536):
        s1 = min(n, s0 + 65536)
        x = anchors[pid[s0:s1]] + noise * torch.randn((s1 - s0, d), device="cuda", generator=g)
        eng.normalize_cast(x, out=out[s0:s1])
    return out, torch.arange(n, device="cuda", dtype=torch.float64) * 0.1 + 1000.0, (pid % 3).to(torch.int32)


for n, d, seg in ((20000, 4096, 62), (20000, 4096, 250), (100000, 1024, 62), (100000, 1024, 250), (300000, 1024, 125)):
    xb, ts, fl = seq(n, d, seg)
    row = {"n": n, "d": d, "seg": seg}
    keys = {}
    for name, symv in (("sym", 1), ("full", -1)):
        p = _native.make_params(k=25, similarity_threshold=0.5, min_time_gap=10.0, max_floor_diff=0, symmetric=symv)
        r = eng.gated_topk(xb, xb, p, q_ts=ts, db_ts=ts, q_floor=fl, db_floor=fl, want_keys=True)
        torch.cuda.synchronize()
        mode = eng.last_sweep_mode()[0]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            r = eng.gated_topk(xb, xb, p, q_ts=ts, db_ts=ts, q_floor=fl, db_floor=fl, want_keys=True)
        e1.record()
        torch.cuda.synchronize()
        row[name + "_ms"] = round(e0.elapsed_time(e1) / 3, 3)
        row[name + "_mode"] = mode
        keys[name] = r.keys
    row["same_lists"] = bool(torch.equal(keys["sym"], keys["full"]))
    row["cand"] = int(r.count.sum().item())
    print(json.dumps(row), flush=True)
    del xb, keys, r
