"""Staged GPU bring-up of the fused kernel with diagnostics (development tool).

    python tools/bringup.py [stage ...]      # default: all stages, each in its own process

Each stage runs in a subprocess with a timeout so that a protocol bug costs one
stage, not the box.  Uses the oracle as the checker (development/test tool only).
"""
from __future__ import annotations

import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "multi-level-indoor-slam_b200"))

import numpy as np  # noqa: E402


def full_matrix_case(cg, Q, N, D, seed=0):
    """N <= 64 and k = 64: the lists hold every score, so the whole Q x N matrix is checked."""
    import torch
    from semgate import _native, synthetic
    from oracle import semgate_oracle as O
    eng = _native.get_engine(0)
    desc = synthetic.make_descriptors(max(Q, N), D, seed=seed)
    q, db = desc[:Q], desc[:N]
    qb = eng.normalize_cast(torch.from_numpy(q).cuda())
    dbb = eng.normalize_cast(torch.from_numpy(db).cuda())
    p = _native.make_params(k=64, cta_group=cg)
    r = eng.gated_topk(qb, dbb, p)
    torch.cuda.synchronize()
    S = np.full((Q, N), np.nan, np.float32)
    idx = r.idx.cpu().numpy()
    sc = r.scores.cpu().numpy()
    ct = r.count.cpu().numpy()
    for i in range(Q):
        S[i, idx[i, :ct[i]]] = sc[i, :ct[i]]
    ref = O.bf16_round(O.l2_normalize(q)).astype(np.float64) @ O.bf16_round(O.l2_normalize(db)).astype(np.float64).T
    err = np.abs(S - ref)
    print(f"[full cg={cg} {Q}x{N}x{D}] counts min/max {ct.min()}/{ct.max()}  nan {np.isnan(S).sum()}  "
          f"max err {np.nanmax(err):.3e}")
    if not (np.nanmax(err) < 1e-4 and not np.isnan(S).any()):
        bad = np.argwhere(~(err < 1e-4))
        print("  first bad entries (row, col, got, want):")
        for rr, cc in bad[:12]:
            print("   ", rr, cc, S[rr, cc], ref[rr, cc])
        print("  bad rows:", np.unique(bad[:, 0])[:40], " bad cols:", np.unique(bad[:, 1])[:40])
        raise SystemExit(1)


def oracle_case(cg, Q, N, D, k, thr=0.5, gap=10.0):
    import torch
    from semgate import _native, synthetic
    from oracle import semgate_oracle as O
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import parity
    eng = _native.get_engine(0)
    desc, ts, fl = synthetic.make_case(max(Q, N), D, 3, seed=Q + N)
    fl = fl.astype(np.int32)
    qb = eng.normalize_cast(torch.from_numpy(desc[:Q]).cuda())
    dbb = eng.normalize_cast(torch.from_numpy(desc[:N]).cuda())
    p = _native.make_params(k=k, similarity_threshold=thr, min_time_gap=gap, max_floor_diff=0, cta_group=cg)
    tts = torch.from_numpy(ts).cuda()
    tfl = torch.from_numpy(fl).cuda()
    r = eng.gated_topk(qb, dbb, p, q_ts=tts[:Q].contiguous(), db_ts=tts[:N].contiguous(),
                       q_floor=tfl[:Q].contiguous(), db_floor=tfl[:N].contiguous())
    torch.cuda.synchronize()
    got = dict(scores=r.scores.cpu().numpy(), idx=r.idx.cpu().numpy().astype(np.int64),
               valid=r.valid.cpu().numpy().astype(bool), count=r.count.cpu().numpy())
    ref = O.gated_topk(desc[:Q], desc[:N], ts[:Q], ts[:N], fl[:Q], fl[:N], k=k, threshold=thr, min_time_gap=gap,
                       max_floor_diff=0, bf16=True)
    rep = parity.compare_candidates(O.compact(ref), O.compact(got), k, thr, tol=1.5e-4)
    print(f"[oracle cg={cg} {Q}x{N}x{D} k={k}] candidates {int(got['count'].sum())} vs {int(ref['count'].sum())}  {rep}")


def timing_case(cg, n, d, k=25, iters=5):
    import torch
    from semgate import _native, synthetic
    eng = _native.get_engine(0)
    x = synthetic.make_descriptors_device(n, d, "cuda", seed=0)
    xb = eng.normalize_cast(x)
    del x
    ts = torch.from_numpy(synthetic.make_timestamps(n)).cuda()
    fl = torch.from_numpy(synthetic.make_floors(n, 3).astype(np.int32)).cuda()
    p = _native.make_params(k=k, similarity_threshold=0.5, min_time_gap=10.0, max_floor_diff=0, cta_group=cg)
    for _ in range(2):
        r = eng.gated_topk(xb, xb, p, q_ts=ts, db_ts=ts, q_floor=fl, db_floor=fl)
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in evs:
        a.record()
        r = eng.gated_topk(xb, xb, p, q_ts=ts, db_ts=ts, q_floor=fl, db_floor=fl)
        b.record()
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in evs)
    flops = 2.0 * n * n * ((d + 63) // 64 * 64)
    best, med = ms[0], ms[len(ms) // 2]
    print(f"[time cg={cg} n={n} d={d}] best {best:.3f} ms  median {med:.3f} ms  "
          f"-> {flops / best / 1e9:.1f} TFLOP/s best, {flops / med / 1e9:.1f} median; "
          f"candidates {int(r.count.sum().item())}")


STAGES = {
    "full1_a": lambda: full_matrix_case(1, 128, 64, 64),
    "full1_b": lambda: full_matrix_case(1, 128, 64, 256),
    "full1_c": lambda: full_matrix_case(1, 100, 50, 1024),
    "full1_d": lambda: full_matrix_case(1, 300, 64, 704),
    "oracle1_a": lambda: oracle_case(1, 300, 700, 128, 25),
    "oracle1_b": lambda: oracle_case(1, 1000, 3000, 512, 25),
    "full2_a": lambda: full_matrix_case(2, 256, 64, 64),
    "full2_b": lambda: full_matrix_case(2, 300, 64, 704),
    "oracle2_a": lambda: oracle_case(2, 300, 700, 128, 25),
    "oracle2_b": lambda: oracle_case(2, 1000, 3000, 512, 25),
    "time1": lambda: timing_case(1, 20000, 4096),
    "time2": lambda: timing_case(2, 20000, 4096),
}


def main():
    names = sys.argv[1:]
    if len(names) == 1 and names[0].startswith("@"):
        STAGES[names[0][1:]]()
        return
    names = names or list(STAGES)
    failed = []
    for n in names:
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "@" + n], timeout=240,
                               capture_output=True, text=True)
            out = (r.stdout + r.stderr).strip().splitlines()
            ok = r.returncode == 0
        except subprocess.TimeoutExpired:
            out, ok = ["TIMEOUT"], False
        print(f"== {n}: {'ok' if ok else 'FAIL'} ({time.time() - t0:.1f}s)")
        for line in out[-25:]:
            print("   " + line)
        sys.stdout.flush()
        if not ok:
            failed.append(n)
            if n.startswith("full1_a"):
                break
    print("FAILED:", failed)
    sys.exit(1 if failed else 0)


if __name__ == "__main__":
    main()
