"""Run under torchrun (one rank per GPU): every multi-GPU form of the sweep must equal the single-GPU sweep
bit for bit -- row-sharded sweeps (merged everywhere / a rank's rows only), triangle-split all-pairs sweeps (forced,
automatic, overflowing -> row-sharded redo, deferred flag), and the host-input forms.  Exit code 0 on success.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/dist_check.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multi-level-indoor-slam_b200"))

import numpy as np
import torch
import torch.distributed as dist

from semgate import _native, synthetic
from semgate.dist import ShardedRetrieval, shard_bounds


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    eng = _native.get_engine(local)
    ok = True
    for exchange in ("allgather", "peer"):
        sr = ShardedRetrieval(eng, exchange=exchange)
        ok = check(sr, eng, dev, rank, world, exchange) and ok
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("dist_check:", "ok" if int(flag.item()) == 1 else "FAILED", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


def same(a, b):
    return (torch.equal(a.idx, b.idx) and torch.equal(a.scores, b.scores) and torch.equal(a.valid, b.valid)
            and torch.equal(a.count, b.count))


def same_rows(part, whole):
    lo, hi = part.lo, part.hi
    r = part.result
    return (torch.equal(r.idx, whole.idx[lo:hi]) and torch.equal(r.scores, whole.scores[lo:hi])
            and torch.equal(r.valid, whole.valid[lo:hi]) and torch.equal(r.count, whole.count[lo:hi]))


def check(sr, eng, dev, rank, world, exchange):
    ok = True
    say = lambda msg: print(f"rank {rank}/{world} [{exchange}]: {msg}", flush=True)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    for (n_db, n_q, d, k, thr, gap) in [(3001, 700, 128, 25, 0.4, 5.0), (20000, 4000, 512, 10, 0.5, 10.0), (9216, 2048, 1024, 25, 0.5, 10.0)]:
        desc, ts, fl = synthetic.make_case(n_db, d, 4, seed=11)
        fl = fl.astype(np.int32)
        lo, hi = shard_bounds(n_db, world, rank)
        xb = eng.normalize_cast(t(desc))
        tts, tfl = t(ts), t(fl)
        mk = lambda off: _native.make_params(k=k, similarity_threshold=thr, min_time_gap=gap, max_floor_diff=0,
                                             db_index_offset=off)
        kw = dict(q_ts=tts[:n_q].contiguous(), db_ts_shard=tts[lo:hi].contiguous(), q_floor=tfl[:n_q].contiguous(),
                  db_floor_shard=tfl[lo:hi].contiguous(), db_floor_all=tfl, max_floor_diff=0)
        whole = eng.gated_topk(xb[:n_q], xb, mk(0), q_ts=tts[:n_q].contiguous(), db_ts=tts, q_floor=tfl[:n_q].contiguous(), db_floor=tfl)
        res = sr.sweep(xb[:n_q], xb[lo:hi], mk, lo, **kw)
        torch.cuda.synchronize()
        good = same(res, whole)
        say(f"n_db={n_db} n_q={n_q} d={d} k={k}: sharded == whole: {good}; candidates {int(whole.count.sum())}")
        ok = ok and good
        # repeated sweeps through the double-buffered symmetric memory (one barrier per step must keep them apart),
        # alternating with the rows-only form
        for it in range(3):
            res2 = sr.sweep(xb[:n_q], xb[lo:hi], mk, lo, **kw)
            part = sr.sweep(xb[:n_q], xb[lo:hi], mk, lo, gather=False, **kw)
            torch.cuda.synchronize()
            good = same(res2, whole) and same_rows(part, whole) and (part.lo, part.hi) == shard_bounds(n_q, world, rank)
            ok = ok and good
        say(f"repeated sweeps + a rank's rows only: ok={good}")
        # host-input form (queries cross PCIe once, on rank 0)
        qh = torch.from_numpy(desc[:n_q].copy()).pin_memory() if rank == 0 else None
        dbh = torch.from_numpy(desc[lo:hi].copy()).pin_memory()
        rh = sr.sweep_from_host(qh, dbh, torch.from_numpy(ts).pin_memory(), torch.from_numpy(fl).pin_memory(), mk, lo, hi, n_q,
                                max_floor_diff=0)
        torch.cuda.synchronize()
        good = same(rh, whole)
        say(f"sweep_from_host == whole: {good}")
        ok = ok and good

        # all-pairs: the ranks split the triangle of tiles (every similarity computed once, on one GPU);
        # with a threshold that admits everything the candidate buffers may overflow -> row-sharded redo
        auto_triangle = n_db >= 8192 and xb.shape[1] >= 1024
        for thr_ap in (thr, -np.inf):
            for force in ((0, 1) if thr_ap == thr else (1,)):
                mk_ap = lambda off: _native.make_params(k=k, similarity_threshold=thr_ap, min_time_gap=gap, max_floor_diff=0,
                                                        db_index_offset=off, symmetric=force)
                p_full = mk_ap(0)
                p_full.symmetric = -1
                full = eng.gated_topk(xb, xb, p_full, q_ts=tts, db_ts=tts, q_floor=tfl, db_floor=tfl)
                ap = sr.sweep_all_pairs(xb, mk_ap, ts=tts, floor=tfl, max_floor_diff=0)
                how = sr.last_all_pairs
                pend = sr.sweep_all_pairs(xb, mk_ap, ts=tts, floor=tfl, max_floor_diff=0, gather=False, defer=True)
                part = pend.result()
                torch.cuda.synchronize()
                good = same(ap, full) and same_rows(part, full)
                if thr_ap == thr:
                    good = good and how == ("triangle" if (force == 1 or auto_triangle) else "rows")
                say(f"all-pairs n={n_db} thr={thr_ap} symmetric={force}: {how}: ok={good}")
                ok = ok and good
        # compacted candidates of a rank's rows carry global query indices; together they are the whole list
        mk1 = lambda off: _native.make_params(k=k, similarity_threshold=thr, min_time_gap=gap, max_floor_diff=0,
                                              db_index_offset=off, symmetric=1)
        p_full = mk1(0)
        p_full.symmetric = -1
        full = eng.gated_topk(xb, xb, p_full, q_ts=tts, db_ts=tts, q_floor=tfl, db_floor=tfl)
        fq, fm, fs, fv, ft = eng.compact(full)
        oq, om, os_, ov, tot = sr.sweep_all_pairs(xb, mk1, ts=tts, floor=tfl, max_floor_diff=0, compact=True, gather=False)
        torch.cuda.synchronize()
        n_tot = int(ft.item())
        plo, phi = shard_bounds(n_db, world, rank)
        sel = (fq[:n_tot] >= plo) & (fq[:n_tot] < phi)
        tt = int(tot.item())
        good = tt == int(sel.sum()) and torch.equal(oq[:tt], fq[:n_tot][sel]) and torch.equal(om[:tt], fm[:n_tot][sel]) \
            and torch.equal(os_[:tt], fs[:n_tot][sel]) and torch.equal(ov[:tt], fv[:n_tot][sel])
        cnt = torch.tensor([tt], device=dev, dtype=torch.int64)
        dist.all_reduce(cnt)
        good = good and int(cnt.item()) == n_tot
        say(f"compacted rows of the triangle sweep: ok={good} ({tt} of {n_tot} candidates here)")
        ok = ok and good
        # host-input all-pairs form (equal row shards only)
        if n_db % world == 0:
            xh = torch.from_numpy(desc[lo:hi].copy()).pin_memory()
            aph = sr.sweep_all_pairs_from_host(xh, torch.from_numpy(ts).pin_memory(), torch.from_numpy(fl).pin_memory(), mk1,
                                               lo, hi, n_db, max_floor_diff=0)
            torch.cuda.synchronize()
            good = same(aph, full)
            say(f"sweep_all_pairs_from_host == full: {good}")
            ok = ok and good
    return ok


if __name__ == "__main__":
    main()
