"""K3 A/B: the dense-list kernel (explicit per-GPU lists: bulk copies, no packing) against the general network kernel.
    python tools/k3_dense_ab.py            # one JSON line per shape"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multi-level-indoor-slam_b200"))

import torch

from semgate import _native

eng = _native.get_engine(0)


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[0], ts[len(ts) // 2]


for (G, Q, k, order) in ((4, 1_000_000, 25, "sorted"), (4, 1_000_000, 25, "unsorted"), (8, 1_000_000, 25, "sorted"),
                         (8, 125_000, 25, "sorted"), (4, 1_000_000, 32, "sorted"), (4, 1_000_000, 10, "sorted"),
                         (4, 1_000_000, 25, "sparse")):
    keys = torch.randint(1, 2 ** 62, (G, Q, k), device="cuda", dtype=torch.int64)
    if order == "sparse":
        keys[torch.rand((G, Q, k), device="cuda") > 0.08] = 0
    if order != "unsorted":
        keys = torch.sort(keys, dim=2, descending=True).values.contiguous()
    byts = G * Q * k * 8 + Q * k * 9 + Q * 4
    row = {"lists": G, "rows": Q, "k": k, "order": order, "algorithmic_bytes": byts}
    outs = {}
    for name, on in (("dense", 1), ("general", 0)):
        eng.set_option("k3_dense", on)
        best, med = timeit(lambda: eng.merge_topk(keys, k))
        row[name + "_ms"] = round(med, 4)
        row[name + "_gbs"] = round(byts / med / 1e6, 1)
        r = eng.merge_topk(keys, k, want_keys=True)
        torch.cuda.synchronize()
        outs[name] = r.keys.clone()
    eng.set_option("k3_dense", 1)
    row["identical"] = bool(torch.equal(outs["dense"], outs["general"]))
    print(json.dumps(row), flush=True)
    del keys, outs
