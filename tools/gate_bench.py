import sys, os, json
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "multi-level-indoor-slam_b200"))
import torch
from semgate import _native
eng = _native.get_engine(0)
def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in evs:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2]
M = 25_000_000
for nl in (2406, 20000, 50 * 1024, 50 * 1024 + 1, 1 << 21):
    fl = torch.randint(1, 6, (nl,), device="cuda", dtype=torch.int32)
    q = (torch.arange(M, device="cuda") // 25 % nl).to(torch.int32)
    m = torch.randint(0, nl, (M,), device="cuda", dtype=torch.int32)
    ms = timeit(lambda: eng.gate_candidates(fl, q, m, 0))
    print(json.dumps({"labels": nl, "pairs": M, "ms": round(ms, 4), "gbs": round(9 * M / ms / 1e6, 1)}), flush=True)
