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


def timing_case(cg, n, d, k=25, iters=5, nq=None):
    """all-pairs when nq is None, else nq queries (the first rows) against n database rows"""
    import torch
    from semgate import _native, synthetic
    eng = _native.get_engine(0)
    dp = (d + 63) // 64 * 64
    xb = torch.empty((n, dp), dtype=torch.bfloat16, device="cuda")
    chunk = max(1, (1 << 28) // (d * 4))
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    places = max(8, n // 20)
    anchors = torch.randn((places, d), generator=g, device="cuda")
    for s0 in range(0, n, chunk):
        e0 = min(n, s0 + chunk)
        pid = torch.randint(0, places, (e0 - s0,), generator=g, device="cuda")
        x = anchors[pid] + 0.6 * torch.randn((e0 - s0, d), generator=g, device="cuda")
        eng.normalize_cast(x, out=xb[s0:e0])
    del x, anchors
    Q = n if nq is None else nq
    qb = xb[:Q]
    ts = torch.from_numpy(synthetic.make_timestamps(n)).cuda()
    fl = torch.from_numpy(synthetic.make_floors(n, 3).astype(np.int32)).cuda()
    p = _native.make_params(k=k, similarity_threshold=0.5, min_time_gap=10.0, max_floor_diff=0, cta_group=cg)
    qts, qfl = ts[:Q].contiguous(), fl[:Q].contiguous()
    eng.set_option("profile", 1)
    for _ in range(2):
        r = eng.gated_topk(qb, xb, p, q_ts=qts, db_ts=ts, q_floor=qfl, db_floor=fl)
    torch.cuda.synchronize()
    eng.profile_read()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in evs:
        a.record()
        r = eng.gated_topk(qb, xb, p, q_ts=qts, db_ts=ts, q_floor=qfl, db_floor=fl)
        b.record()
    torch.cuda.synchronize()
    k2ms, k2n = eng.profile_read()
    ms = sorted(a.elapsed_time(b) for a, b in evs)
    flops = 2.0 * Q * n * dp
    best, med = ms[0], ms[len(ms) // 2]
    print(f"[time cg={cg} Q={Q} N={n} d={d}] best {best:.3f} ms  median {med:.3f} ms  K2 avg {k2ms / k2n:.3f} ms "
          f"-> {flops / (k2ms / k2n) / 1e9:.1f} TFLOP/s (K2), {flops / med / 1e9:.1f} (K2+K3 median); "
          f"pairs/s {Q * n / med * 1e3:.3e}; candidates {int(r.count.sum().item())}")


def ab_case(n, d, nq=None, rounds=8, k=25, sustain_s=0.0):
    """Interleaved A/B of cta_group 1 vs 2 in one process (same thermal state), K2-only times."""
    import torch
    from semgate import _native, synthetic
    eng = _native.get_engine(0)
    dp = (d + 63) // 64 * 64
    xb = torch.empty((n, dp), dtype=torch.bfloat16, device="cuda")
    chunk = max(1, (1 << 28) // (d * 4))
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    places = max(8, n // 20)
    anchors = torch.randn((places, d), generator=g, device="cuda")
    for s0 in range(0, n, chunk):
        e0 = min(n, s0 + chunk)
        pid = torch.randint(0, places, (e0 - s0,), generator=g, device="cuda")
        x = anchors[pid] + 0.6 * torch.randn((e0 - s0, d), generator=g, device="cuda")
        eng.normalize_cast(x, out=xb[s0:e0])
    del x, anchors
    Q = n if nq is None else nq
    qb = xb[:Q]
    ts = torch.from_numpy(synthetic.make_timestamps(n)).cuda()
    fl = torch.from_numpy(synthetic.make_floors(n, 3).astype(np.int32)).cuda()
    qts, qfl = ts[:Q].contiguous(), fl[:Q].contiguous()
    eng.set_option("profile", 1)
    flops = 2.0 * Q * n * dp
    ps = {cg: _native.make_params(k=k, similarity_threshold=0.5, min_time_gap=10.0, max_floor_diff=0, cta_group=cg)
          for cg in (1, 2)}
    times = {1: [], 2: []}
    for r in range(rounds + 1):
        for cg in (1, 2):
            reps = 3
            if sustain_s > 0:
                reps = max(3, int(sustain_s / max(times[cg][-1] * 1e-3, 1e-4))) if times[cg] else 3
            for _ in range(reps):
                eng.gated_topk(qb, xb, ps[cg], q_ts=qts, db_ts=ts, q_floor=qfl, db_floor=fl)
            torch.cuda.synchronize()
            ms, nn = eng.profile_read()
            if r > 0 or sustain_s == 0:
                times[cg].append(ms / nn)
            elif not times[cg]:
                times[cg].append(ms / nn)
    for cg in (1, 2):
        t = sorted(times[cg][1:])
        print(f"[ab cg={cg} Q={Q} N={n} d={d} sustain={sustain_s}] K2 min {t[0]:.3f} med {t[len(t) // 2]:.3f} max {t[-1]:.3f} ms"
              f" -> {flops / t[0] / 1e9:.0f} / {flops / t[len(t) // 2] / 1e9:.0f} / {flops / t[-1] / 1e9:.0f} TFLOP/s")


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
    "time1_c1": lambda: timing_case(1, 5000, 512),
    "time2_c1": lambda: timing_case(2, 5000, 512),
    "time1_c3": lambda: timing_case(1, 100000, 8448, nq=10000),
    "time2_c3": lambda: timing_case(2, 100000, 8448, nq=10000),
    "ab_c2": lambda: ab_case(20000, 4096),
    "ab_c2_sustained": lambda: ab_case(20000, 4096, rounds=4, sustain_s=1.5),
    "ab_c3": lambda: ab_case(100000, 8448, nq=10000, rounds=5),
    "ab_50k": lambda: ab_case(50000, 4096, rounds=5),
    "time1_q1": lambda: timing_case(1, 100000, 4096, nq=1, iters=20),
    "time1_100k": lambda: timing_case(1, 100000, 4096, iters=3),
    "time2_100k": lambda: timing_case(2, 100000, 4096, iters=3),
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
