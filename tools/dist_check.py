"""Run under torchrun (one rank per GPU): row-sharded sweep + NCCL all-gather + merge must equal
the single-GPU sweep bit for bit.  Exit code 0 on success.

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
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


def check(sr, eng, dev, rank, world, exchange):
    ok = True
    for (n_db, n_q, d, k, thr, gap) in [(3001, 700, 128, 25, 0.4, 5.0), (20000, 4000, 512, 10, 0.5, 10.0)]:
        desc, ts, fl = synthetic.make_case(n_db, d, 4, seed=11)
        fl = fl.astype(np.int32)
        lo, hi = shard_bounds(n_db, world, rank)
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        xb = eng.normalize_cast(t(desc))
        tts, tfl = t(ts), t(fl)
        mk = lambda off: _native.make_params(k=k, similarity_threshold=thr, min_time_gap=gap, max_floor_diff=0,
                                             db_index_offset=off)
        res = sr.sweep(xb[:n_q], xb[lo:hi], mk, lo, q_ts=tts[:n_q].contiguous(), db_ts_shard=tts[lo:hi].contiguous(),
                       q_floor=tfl[:n_q].contiguous(), db_floor_shard=tfl[lo:hi].contiguous(), db_floor_all=tfl,
                       max_floor_diff=0)
        whole = eng.gated_topk(xb[:n_q], xb, mk(0), q_ts=tts[:n_q].contiguous(), db_ts=tts,
                               q_floor=tfl[:n_q].contiguous(), db_floor=tfl)
        torch.cuda.synchronize()
        same = (torch.equal(res.idx, whole.idx) and torch.equal(res.scores, whole.scores)
                and torch.equal(res.valid, whole.valid) and torch.equal(res.count, whole.count))
        print(f"rank {rank}/{world} [{exchange}]: n_db={n_db} n_q={n_q} d={d} k={k}: sharded == whole: {same}; "
              f"candidates {int(whole.count.sum())}", flush=True)
        ok = ok and same
        # a second sweep through the same symmetric buffer (the barriers must keep steps apart)
        res2 = sr.sweep(xb[:n_q], xb[lo:hi], mk, lo, q_ts=tts[:n_q].contiguous(), db_ts_shard=tts[lo:hi].contiguous(),
                        q_floor=tfl[:n_q].contiguous(), db_floor_shard=tfl[lo:hi].contiguous(), db_floor_all=tfl,
                        max_floor_diff=0)
        torch.cuda.synchronize()
        ok = ok and torch.equal(res2.idx, whole.idx) and torch.equal(res2.scores, whole.scores)
        # all-pairs: the ranks split the triangle of tiles (every similarity computed once, on one GPU);
        # with a threshold that admits everything the candidate buffers may overflow -> row-sharded redo
        for thr_ap in (thr, -np.inf):
            mk_ap = lambda off: _native.make_params(k=k, similarity_threshold=thr_ap, min_time_gap=gap, max_floor_diff=0,
                                                    db_index_offset=off)
            p_full = mk_ap(0)
            p_full.symmetric = -1
            full = eng.gated_topk(xb, xb, p_full, q_ts=tts, db_ts=tts, q_floor=tfl, db_floor=tfl)
            ap = sr.sweep_all_pairs(xb, mk_ap, ts=tts, floor=tfl, max_floor_diff=0)
            torch.cuda.synchronize()
            good = (torch.equal(ap.idx, full.idx) and torch.equal(ap.scores, full.scores)
                    and torch.equal(ap.valid, full.valid) and torch.equal(ap.count, full.count))
            if thr_ap == thr:
                good = good and sr.last_all_pairs == "triangle"
            print(f"rank {rank}/{world} [{exchange}]: all-pairs n={n_db} thr={thr_ap}: {sr.last_all_pairs}: ok={good}", flush=True)
            ok = ok and good
    return ok


if __name__ == "__main__":
    main()
