"""Host-buffer call at config 2: timestamps / labels pageable against pinned, three rounds of 10 calls each (the first round also
shows the warm-up the call needs after idle).  python tools/e2e_pin_ab.py"""
import sys, os, time, json
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "multi-level-indoor-slam_b200"))
import numpy as np, torch
from semgate import _native, synthetic
eng = _native.get_engine(0)
n, d = 20000, 4096
desc, ts, fl = synthetic.make_case(n, d, 3, seed=0)
fl = fl.astype(np.int32)
p = _native.make_params(k=25, similarity_threshold=0.5, min_time_gap=10.0, max_floor_diff=0)
q = torch.from_numpy(desc).pin_memory(); qh = q.numpy()
cap = n * 25
outs = tuple(torch.empty((cap,), dtype=dt, pin_memory=True).numpy() for dt in (torch.int32, torch.int32, torch.float32, torch.uint8))
tsp_t, flp_t = torch.from_numpy(ts).pin_memory(), torch.from_numpy(fl).pin_memory()
tsp, flp = tsp_t.numpy(), flp_t.numpy()
def run(a, b, reps=10):
    eng.find_loop_closures_host(qh, a, b, p, out=outs)
    t0 = time.perf_counter()
    for _ in range(reps):
        r = eng.find_loop_closures_host(qh, a, b, p, out=outs)
    return (time.perf_counter() - t0) / reps * 1e3, len(r[0])
for i in range(3):
    print("pageable ts/fl", run(ts, fl), "pinned ts/fl", run(tsp, flp), flush=True)
