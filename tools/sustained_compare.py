"""How do clocks / throughput evolve over ~20 s of back-to-back work: cuBLAS bf16 GEMM vs the fused sweep.
(development tool; answers whether a long sweep is throttled harder than a library GEMM)"""
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multi-level-indoor-slam_b200"))

import numpy as np
import torch

from semgate import _native, synthetic


def sample_clock():
    r = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,temperature.gpu", "--format=csv,noheader,nounits"],
                       capture_output=True, text=True)
    return r.stdout.strip()


def run(name, fn, flops, seconds=16.0, window=2.0):
    torch.cuda.synchronize()
    t_end = time.time() + seconds
    while time.time() < t_end:
        t0 = time.time()
        n = 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        while time.time() - t0 < window:
            for _ in range(8):
                fn()
            n += 8
            torch.cuda.synchronize()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(f"{name}: {flops * n / ms / 1e9:7.0f} TFLOP/s over {ms / 1e3:.1f} s   [sm MHz, W, C] = {sample_clock()}", flush=True)


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "both"
    if which in ("cublas", "both"):
        a = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
        b = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
        run("cublas 8192^3", lambda: torch.matmul(a, b), 2.0 * 8192 ** 3)
        del a, b
        time.sleep(5)
    if which in ("ours", "both"):
        eng = _native.get_engine(0)
        n, d = 20000, 4096
        xb = eng.normalize_cast(synthetic.make_descriptors_device(n, d, "cuda", seed=0))
        ts = torch.from_numpy(synthetic.make_timestamps(n)).cuda()
        fl = torch.from_numpy(synthetic.make_floors(n, 3).astype(np.int32)).cuda()
        for cg in (1, 2):
            p = _native.make_params(k=25, similarity_threshold=0.5, min_time_gap=10.0, max_floor_diff=0, cta_group=cg)
            run(f"semgate cg{cg} 20k x 20k x 4096", lambda: eng.gated_topk(xb, xb, p, q_ts=ts, db_ts=ts, q_floor=fl, db_floor=fl),
                2.0 * n * n * d)
            time.sleep(5)
    if which in ("randn",):
        # same sweep on unstructured gaussian rows (toggle-rate comparison)
        eng = _native.get_engine(0)
        n, d = 20000, 4096
        xb = eng.normalize_cast(torch.randn(n, d, device="cuda"))
        p = _native.make_params(k=25, similarity_threshold=0.5, cta_group=2)
        run("semgate cg2 randn rows", lambda: eng.gated_topk(xb, xb, p), 2.0 * n * n * d)


if __name__ == "__main__":
    main()
