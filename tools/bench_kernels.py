"""CUDA-event timings of the HBM-bound companion kernels against the measured copy bandwidth.

    python tools/bench_kernels.py > gpurun_out/kernels.json
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multi-level-indoor-slam_b200"))

import numpy as np
import torch

from semgate import _native, synthetic


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in evs:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    t = sorted(a.elapsed_time(b) for a, b in evs)
    return t[0], t[len(t) // 2]


def main():
    eng = _native.get_engine(0)
    peak = 6543.1
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        peak = float(json.load(open(p))["hbm_gbs"])
    out = {"hbm_peak_gbs": peak, "kernels": []}

    # K1: normalise + cast, larger than L2
    for n, d in ((65536, 4096), (32768, 8448), (8192, 49152)):
        x = torch.randn((n, d), device="cuda")
        o = torch.empty((n, _native.pad_dim(d)), dtype=torch.bfloat16, device="cuda")
        best, med = timeit(lambda: eng.normalize_cast(x, out=o))
        byts = n * (4 * d + 2 * _native.pad_dim(d))
        out["kernels"].append({"kernel": "K1 normalize_cast", "shape": [n, d], "ms": med, "algorithmic_bytes": byts,
                               "gbs": byts / med / 1e6, "frac_of_hbm_peak": byts / med / 1e6 / peak})
        del x, o

    # streaming query(): one query row against a large resident database (HBM-bound GEMV shape)
    for n, d in ((200000, 4096), (100000, 8448)):
        dp = _native.pad_dim(d)
        db = torch.randn((n, dp), device="cuda", dtype=torch.bfloat16)
        q = db[:1].contiguous()
        prm = _native.make_params(k=25)
        best, med = timeit(lambda: eng.gated_topk(q, db, prm), iters=20)
        byts = 2 * n * dp
        out["kernels"].append({"kernel": "K6+K3, streaming single query (query())", "shape": [1, n, d], "ms": med,
                               "algorithmic_bytes": byts, "gbs": byts / med / 1e6, "frac_of_hbm_peak": byts / med / 1e6 / peak,
                               "queries_per_s": 1e3 / med})
        del db

    # gate over explicit pairs
    M, nl = 50_000_000, 1_000_000
    fl = torch.randint(1, 6, (nl,), device="cuda", dtype=torch.int32)
    qi = torch.randint(0, nl, (M,), device="cuda", dtype=torch.int32)
    mi = torch.randint(0, nl, (M,), device="cuda", dtype=torch.int32)
    best, med = timeit(lambda: eng.gate_candidates(fl, qi, mi, 0))
    byts = 9 * M
    out["kernels"].append({"kernel": "gate_candidates", "shape": [M], "ms": med, "algorithmic_bytes": byts,
                           "gbs": byts / med / 1e6, "frac_of_hbm_peak": byts / med / 1e6 / peak, "candidates_per_s": M / med * 1e3})
    # the order K4 emits: query index ascending, k = 25 matches each (label gathers of the query side coalesce)
    qi = torch.arange(M // 25, device="cuda", dtype=torch.int32).repeat_interleave(25) % nl
    best, med = timeit(lambda: eng.gate_candidates(fl, qi, mi, 0))
    out["kernels"].append({"kernel": "gate_candidates (query-sorted pairs, as compacted)", "shape": [M], "ms": med,
                           "algorithmic_bytes": byts, "gbs": byts / med / 1e6, "frac_of_hbm_peak": byts / med / 1e6 / peak,
                           "candidates_per_s": M / med * 1e3})
    del fl, qi, mi

    # K3 + K4 on a 1M-row result
    Q, k = 1_000_000, 25
    x = synthetic.make_descriptors_device(4096, 256, "cuda")
    xb = eng.normalize_cast(x)
    keys = torch.randint(1, 2**62, (4, Q, k), device="cuda", dtype=torch.int64)
    best, med = timeit(lambda: eng.merge_topk(keys, k))
    byts = Q * k * (8 * 4 + 9) + 4 * Q
    out["kernels"].append({"kernel": "K3 merge_topk (4 lists)", "shape": [4, Q, k], "ms": med, "algorithmic_bytes": byts,
                           "gbs": byts / med / 1e6, "frac_of_hbm_peak": byts / med / 1e6 / peak})
    res = eng.merge_topk(keys, k)
    best, med = timeit(lambda: eng.compact(res))
    byts = Q * k * (9 + 13) + 4 * Q
    out["kernels"].append({"kernel": "K4 compact", "shape": [Q, k], "ms": med, "algorithmic_bytes": byts,
                           "gbs": byts / med / 1e6, "frac_of_hbm_peak": byts / med / 1e6 / peak})
    del keys, res
    # K5: CricaVPR cross-correlation re-rank, DINOv2 shape (529 patches x 768-d), 25 candidates per query
    nf, P, Dl, kc = 2000, 529, 768, 25
    feats = torch.empty((nf, P, _native.pad_dim(Dl)), dtype=torch.bfloat16, device="cuda")
    for s0 in range(0, nf, 250):
        x = torch.randn((250 * P, Dl), device="cuda")
        eng.normalize_cast(x, out=feats[s0:s0 + 250].view(250 * P, -1))
    nq = 4000
    qi = torch.arange(nq, device="cuda", dtype=torch.int32).repeat_interleave(kc) % nf
    mi = torch.randint(0, nf, (nq * kc,), device="cuda", dtype=torch.int32)
    gs = torch.rand((nq * kc,), device="cuda")
    best, med = timeit(lambda: eng.rerank_scores(feats, qi, mi, gs), iters=5, warm=2)
    flops = 2.0 * P * P * _native.pad_dim(Dl) * nq * kc
    out["kernels"].append({"kernel": "K5 rerank_scores (cross-correlation, 529 x 768 patches)", "shape": [nq * kc, P, Dl], "ms": med,
                           "pairs_per_s": nq * kc / med * 1e3, "tflops_useful": flops / med / 1e9,
                           "algorithmic_bytes": 0, "gbs": 0.0, "frac_of_hbm_peak": 0.0,
                           "note": "useful FLOPs 2*P*P*D per pair; the 128 x 256 tiling computes 640 x 544 per pair"})
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
