"""Schedule / L2-policy A-B sweep of the fused kernel (development tool).

    python tools/l2_sweep.py 100000,300000 "cg,cap,window,hint;cg,cap,window,hint;..." [seconds]

Each configuration is applied through the environment knobs the library reads per launch
(SEMGATE_RM_CAP_MB, SEMGATE_WINDOW_MB, SEMGATE_L2_HINT) and timed back to back for `seconds`
with NVML clock / power samples taken WHILE the kernel runs.
"""
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multi-level-indoor-slam_b200"))

import numpy as np
import torch

from semgate import _native, synthetic


class Nvml:
    def __init__(self):
        import pynvml
        self.n = pynvml
        pynvml.nvmlInit()
        self.h = pynvml.nvmlDeviceGetHandleByIndex(0)
        self.samples = []
        self.run = False

    def _loop(self):
        while self.run:
            try:
                self.samples.append((self.n.nvmlDeviceGetClockInfo(self.h, self.n.NVML_CLOCK_SM),
                                     self.n.nvmlDeviceGetPowerUsage(self.h) / 1000.0))
            except Exception:
                pass
            time.sleep(0.02)

    def start(self):
        self.samples, self.run = [], True
        self.t = threading.Thread(target=self._loop, daemon=True)
        self.t.start()

    def stop(self):
        self.run = False
        self.t.join()
        if not self.samples:
            return 0, 0
        a = np.array(self.samples)
        return float(np.median(a[:, 0])), float(np.median(a[:, 1]))


def main():
    # sizes: N (all pairs) or QxN
    sizes = [tuple(int(v) for v in x.split("x")) if "x" in x else (int(x), int(x)) for x in sys.argv[1].split(",")]
    cfgs = [tuple(int(v) for v in c.split(",")) for c in sys.argv[2].split(";")]
    secs = float(sys.argv[3]) if len(sys.argv) > 3 else 2.5
    d = int(os.environ.get("SWEEP_DIM", "4096"))
    eng = _native.get_engine(0)
    eng.set_option("profile", 1)
    nv = Nvml()
    nmax = max(max(a, b) for a, b in sizes)
    xb = torch.empty((nmax, d), dtype=torch.bfloat16, device="cuda")
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    places = max(8, nmax // 20)
    anchors = torch.randn((places, d), generator=g, device="cuda")
    step = max(1024, (1 << 27) // d)
    for s0 in range(0, nmax, step):
        e0 = min(nmax, s0 + step)
        pid = torch.randint(0, places, (e0 - s0,), generator=g, device="cuda")
        eng.normalize_cast(anchors[pid] + 0.6 * torch.randn((e0 - s0, d), generator=g, device="cuda"), out=xb[s0:e0])
    del anchors
    ts = torch.from_numpy(synthetic.make_timestamps(nmax)).cuda()
    fl = torch.from_numpy(synthetic.make_floors(nmax, 16).astype(np.int32)).cuda()
    ref_idx = {}
    for nq, n in sizes:
        q, t, f = xb[:n], ts[:n].contiguous(), fl[:n].contiguous()
        qq, tq, fq = xb[:nq], ts[:nq].contiguous(), fl[:nq].contiguous()
        for cg, cap, window, hint in cfgs:
            os.environ["SEMGATE_RM_CAP_MB"] = str(cap)
            os.environ["SEMGATE_WINDOW_MB"] = str(window)
            os.environ["SEMGATE_L2_HINT"] = str(hint)
            p = _native.make_params(k=25, similarity_threshold=0.5, min_time_gap=10.0, max_floor_diff=0, cta_group=cg)
            r = eng.gated_topk(qq, q, p, q_ts=tq, db_ts=t, q_floor=fq, db_floor=f)
            torch.cuda.synchronize()
            chk = int(r.idx.to(torch.int64).sum().item())
            same = ref_idx.setdefault((nq, n), chk) == chk
            eng.profile_read()
            nv.start()
            t0 = time.time()
            while True:
                eng.gated_topk(qq, q, p, q_ts=tq, db_ts=t, q_floor=fq, db_floor=f)
                torch.cuda.synchronize()
                if time.time() - t0 >= secs:
                    break
            mhz, watts = nv.stop()
            ms, nn = eng.profile_read()
            print(f"n={nq}x{n}x{d} cg={cg} cap={cap} window_mb={window} hint={hint}: {2.0 * nq * n * d * nn / ms / 1e9:7.0f} TFLOP/s "
                  f"({ms / nn:9.2f} ms/launch, {nn} launches)  sm {mhz:.0f} MHz  {watts:.0f} W  same_result={same}", flush=True)
            time.sleep(1.0)


if __name__ == "__main__":
    main()
