"""Throughput / clock / power of the fused sweep as the all-pairs problem grows (development tool)."""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multi-level-indoor-slam_b200"))

import numpy as np
import torch

from semgate import _native, synthetic


def smi():
    r = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,power.draw,temperature.gpu,temperature.memory",
                        "--format=csv,noheader,nounits"], capture_output=True, text=True)
    return r.stdout.strip()


def main():
    sizes = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [20000, 100000, 300000]
    cgs = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [2]
    secs = float(sys.argv[3]) if len(sys.argv) > 3 else 8.0
    thr = float(sys.argv[4]) if len(sys.argv) > 4 else 0.5
    eng = _native.get_engine(0)
    eng.set_option("profile", 1)
    d = 4096
    nmax = max(sizes)
    xb = torch.empty((nmax, d), dtype=torch.bfloat16, device="cuda")
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    places = max(8, nmax // 20)
    anchors = torch.randn((places, d), generator=g, device="cuda")
    for s0 in range(0, nmax, 32768):
        e0 = min(nmax, s0 + 32768)
        pid = torch.randint(0, places, (e0 - s0,), generator=g, device="cuda")
        eng.normalize_cast(anchors[pid] + 0.6 * torch.randn((e0 - s0, d), generator=g, device="cuda"), out=xb[s0:e0])
    del anchors
    ts = torch.from_numpy(synthetic.make_timestamps(nmax)).cuda()
    fl = torch.from_numpy(synthetic.make_floors(nmax, 16).astype(np.int32)).cuda()
    for n in sizes:
        for cg in cgs:
            p = _native.make_params(k=25, similarity_threshold=thr, min_time_gap=10.0, max_floor_diff=0, cta_group=cg)
            q, t, f = xb[:n], ts[:n].contiguous(), fl[:n].contiguous()
            t_end = time.time() + secs
            eng.gated_topk(q, q, p, q_ts=t, db_ts=t, q_floor=f, db_floor=f)
            torch.cuda.synchronize()
            eng.profile_read()
            while time.time() < t_end:
                t0 = time.time()
                while time.time() - t0 < 2.0:
                    eng.gated_topk(q, q, p, q_ts=t, db_ts=t, q_floor=f, db_floor=f)
                    torch.cuda.synchronize()
                ms, nn = eng.profile_read()
                print(f"n={n} cg={cg}: {2.0 * n * n * d * nn / ms / 1e9:7.0f} TFLOP/s ({ms / nn:9.2f} ms/launch)  "
                      f"[sm MHz, mem MHz, W, C gpu, C mem] = {smi()}", flush=True)
            time.sleep(3)


if __name__ == "__main__":
    main()
