"""K4 (candidate compaction) at three densities: 1M x 25 full lists (the companion benchmark), 1M rows with ~20 and ~1
candidates each, 20k rows (config 2's shape).  One JSON line per case; checks the flat list against torch."""
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
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in evs:                      # queued back to back: the host's launch latency stays outside
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2]


CASES = ((1_000_000, 25, 1.0), (1_000_000, 25, 0.8), (1_000_000, 25, 0.04), (20_000, 25, 0.8), (1_000_000, 10, 1.0),
         (300_000, 64, 0.5))
if len(sys.argv) > 1:
    CASES = CASES[:int(sys.argv[1])]
for Q, k, fill in CASES:
    keys = torch.randint(1, 2 ** 62, (1, Q, k), device="cuda", dtype=torch.int64)
    if fill < 1.0:
        keys[torch.rand((1, Q, k), device="cuda") > fill] = 0
    res = eng.merge_topk(keys, k)
    del keys
    ms = timeit(lambda: eng.compact(res))
    oq, om, os_, ov, tot = eng.compact(res)
    M = int(tot.item())
    sel = torch.arange(k, device="cuda")[None, :] < res.count[:, None]
    rows = torch.arange(Q, device="cuda", dtype=torch.int32)[:, None].expand(Q, k)
    ok = (M == int(sel.sum().item()) and torch.equal(oq[:M], rows[sel]) and torch.equal(om[:M], res.idx[sel])
          and torch.equal(os_[:M].view(torch.int32), res.scores[sel].view(torch.int32)) and torch.equal(ov[:M], res.valid[sel]))
    byts = Q * k * 9 + 4 * Q + 13 * M
    print(json.dumps({"variant": os.environ.get("SEMGATE_K4_VARIANT", "default"), "rows": Q, "k": k, "fill": fill, "candidates": M, "ms": round(ms, 4), "gbs": round(byts / ms / 1e6, 1),
                      "bytes": byts, "correct": bool(ok)}), flush=True)
    del res, oq, om, os_, ov, sel, rows
