"""Short profiling target for K3: 1M rows x 4 full lists of 25 keys (the companion-kernel benchmark shape)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multi-level-indoor-slam_b200"))

import torch

from semgate import _native

eng = _native.get_engine(0)
keys = torch.randint(1, 2 ** 62, (4, 1_000_000, 25), device="cuda", dtype=torch.int64)
for _ in range(4):
    r = eng.merge_topk(keys, 25)
torch.cuda.synchronize()
print("rows with a full list", int((r.count == 25).sum().item()))
