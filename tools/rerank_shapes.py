"""K5 pair form over patch counts that select different tilings (tile width, strip or not): executed and useful TFLOP/s.
    python tools/rerank_shapes.py"""
import os, sys, json, torch
sys.path.insert(0, "multi-level-indoor-slam_b200")
from semgate import _native
eng = _native.get_engine(0)
def executed(P):
    full, rem = P // 256, P % 256
    strip = full >= 1 and 0 < rem <= 32
    bnmax = 192 if strip else 256
    mt2 = full if strip else (P + 255) // 256
    nt = max((P + bnmax - 1) // bnmax, 2 if mt2 == 1 else 1)
    bn = (((P + nt - 1) // nt) + 31) // 32 * 32
    ncols = sum(((min(bn, P - t * bn) + 15) // 16) * 16 for t in range(nt))
    return (mt2 * 256 * ncols + (ncols * 32 if strip else 0)), bn, nt, mt2, strip
for P in (512, 576, 529, 768, 384, 640):
    nf, Dl, kc, nq = 600, 768, 25, 2000
    feats = torch.empty((nf, P, _native.pad_dim(Dl)), dtype=torch.bfloat16, device="cuda")
    for s0 in range(0, nf, 200):
        x = torch.randn((200 * P, Dl), device="cuda")
        eng.normalize_cast(x, out=feats[s0:s0 + 200].view(200 * P, -1))
    qi = torch.arange(nq, device="cuda", dtype=torch.int32).repeat_interleave(kc) % nf
    mi = torch.randint(0, nf, (nq * kc,), device="cuda", dtype=torch.int32)
    gs = torch.rand((nq * kc,), device="cuda")
    for _ in range(2): eng.rerank_scores(feats, qi, mi, gs)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eng.rerank_scores(feats, qi, mi, gs); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[2]
    mm, bn, nt, mt2, strip = executed(P)
    print(json.dumps({"P": P, "bn": bn, "nt": nt, "mt2": mt2, "strip": strip, "ms": round(ms, 3), "us_per_pair_per_cluster": round(ms * 1e3 / (nq * kc / 74), 2),
                      "tflops_executed": round(2.0 * mm * Dl * nq * kc / ms / 1e9, 1), "tflops_useful": round(2.0 * P * P * Dl * nq * kc / ms / 1e9, 1)}), flush=True)
    del feats
