"""Profiling target for K5: DINOv2-shape re-rank of `pairs` candidate pairs (25 per query).

    ncu --set full -k regex:rerank_kernel -c 1 python tools/ncu_rerank.py [pairs]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multi-level-indoor-slam_b200"))

import torch

from semgate import _native

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
eng = _native.get_engine(0)
nf, P, Dl, kc = 1000, 529, 768, 25
feats = torch.empty((nf, P, _native.pad_dim(Dl)), dtype=torch.bfloat16, device="cuda")
for s0 in range(0, nf, 250):
    x = torch.randn((250 * P, Dl), device="cuda")
    eng.normalize_cast(x, out=feats[s0:s0 + 250].view(250 * P, -1))
nq = pairs // kc
qi = torch.arange(nq, device="cuda", dtype=torch.int32).repeat_interleave(kc) % nf
mi = torch.randint(0, nf, (nq * kc,), device="cuda", dtype=torch.int32)
gs = torch.rand((nq * kc,), device="cuda")
for _ in range(2):
    cross, comb = eng.rerank_scores(feats, qi, mi, gs)
torch.cuda.synchronize()
print("mean combined", float(comb.mean()))
