"""End-to-end host-buffer call (semgate_find_loop_closures_host) at config 2 under different row-chunk schemes
(SEMGATE_E2E_CHUNKS: comma-separated weights, first chunk first).     python tools/e2e_chunks.py"""
import os, sys, time, json, numpy as np, torch
sys.path.insert(0, "multi-level-indoor-slam_b200")
from semgate import _native, synthetic
eng = _native.get_engine(0)
n = 20000
desc, ts, fl = synthetic.make_case(n, 4096, 3, seed=0)
fl = fl.astype(np.int32)
qh = torch.from_numpy(desc).pin_memory().numpy()
cap = n * 25
outs = tuple(torch.empty((cap,), dtype=dt, pin_memory=True).numpy() for dt in (torch.int32, torch.int32, torch.float32, torch.uint8))
p = _native.make_params(k=25, similarity_threshold=0.5, min_time_gap=10.0, max_floor_diff=0)
def run(tag, env):
    for k, v in env.items():
        if v: os.environ[k] = v
        else: os.environ.pop(k, None)
    r = eng.find_loop_closures_host(qh, ts, fl, p, out=outs)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10): r = eng.find_loop_closures_host(qh, ts, fl, p, out=outs)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / 10 * 1e3, len(r[0])
schemes = {"default": "", "r1": "4,4,4,4,3,3,2,2,1,1", "tail2": "4,4,4,4,3,3,2,2,2", "equal14": "2,2,2,2,2,2,2,2,2,2,2,2,2,2", "five": "5,4,3,2,1"}
for rep in range(2):
    for tag, w in schemes.items():
        ms, c = run(tag, {"SEMGATE_E2E_CHUNKS": w})
        print(rep, tag, w, round(ms, 3), c, flush=True)
