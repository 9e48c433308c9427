"""Short profiling target for the HBM-bound companions at the benchmark's shapes: K3 (dense-list kernel, 1M rows x 4 full
sorted lists of 25), K4 (1M x 25 full lists), the pair gate over K4's output (2M labels: global gathers; 20 000 labels: table in shared memory).  Three rounds; capture the third
(ncu -k regex:"merge_dense|compact_onepass|gate_candidates" -s 8 -c 4)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multi-level-indoor-slam_b200"))

import torch

from semgate import _native

eng = _native.get_engine(0)
Q, k = 1_000_000, 25
keys = torch.sort(torch.randint(1, 2 ** 62, (4, Q, k), device="cuda", dtype=torch.int64), dim=2, descending=True).values.contiguous()
fl = torch.randint(1, 6, (1 << 21,), device="cuda", dtype=torch.int32)
for _ in range(3):
    res = eng.merge_topk(keys, k)
    oq, om, os_, ov, tot = eng.compact(res)
    M = Q * k
    eng.gate_candidates(fl, oq[:M] % fl.shape[0], om[:M] % fl.shape[0], 0)
    eng.gate_candidates(fl[:20000].contiguous(), oq[:M] % 20000, om[:M] % 20000, 0)      # label table in shared memory
torch.cuda.synchronize()
print("candidates", int(tot.item()))
