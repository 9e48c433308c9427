"""Short profiling target: the BASELINE config-2 sweep (20k x 20k x 4096-d, k=25) a few times.

    python tools/ncu_target.py [cta_group] [n] [d] [iters]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multi-level-indoor-slam_b200"))

import numpy as np
import torch

from semgate import _native, synthetic

cg = int(sys.argv[1]) if len(sys.argv) > 1 else 1
n = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
d = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 4

eng = _native.get_engine(0)
x = synthetic.make_descriptors_device(n, d, "cuda", seed=0)
xb = eng.normalize_cast(x)
del x
ts = torch.from_numpy(synthetic.make_timestamps(n)).cuda()
fl = torch.from_numpy(synthetic.make_floors(n, 3).astype(np.int32)).cuda()
p = _native.make_params(k=25, similarity_threshold=0.5, min_time_gap=10.0, max_floor_diff=0, cta_group=cg)
for _ in range(iters):
    r = eng.gated_topk(xb, xb, p, q_ts=ts, db_ts=ts, q_floor=fl, db_floor=fl)
    out = eng.compact(r)
torch.cuda.synchronize()
print("candidates", int(out[4].item()))
