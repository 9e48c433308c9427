"""CPU tier: the row-sharded multi-rank path (semgate/dist.py) with world_size 2 over gloo.

The orchestration under test is the product's: shard bounds, global column offsets, the
all-gather of per-rank candidate keys and the merge call.  The two kernels it drives
(K2 sweep, K3 merge) need a GPU, so a stand-in engine restates them with the oracle on
CPU tensors; the same merge is checked against the real K3 kernel in tests/test_gpu_parity.py.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import semgate_oracle as O
from semgate import synthetic
from semgate.dist import ShardedRetrieval, shard_bounds


class _Res:
    def __init__(self, **kw):
        self.__dict__.update(kw)


class OracleEngine:
    """Stand-in for semgate._native.Engine on CPU tensors (test double)."""

    torch_device = torch.device("cpu")
    device = 0

    def __init__(self, desc_q, desc_db_all):
        self.q, self.db_all = desc_q, desc_db_all

    def normalize_cast(self, x, out=None):
        return x          # the stand-in sweep normalises inside the oracle call

    overflow_on_rank = None       # test knob: this rank reports an overflowed symmetric part
    _mode = 0

    def last_sweep_mode(self):
        return self._mode, 0

    def last_sweep_overflow(self, out=None):
        return torch.tensor([1 if self._mode == 2 else 0], dtype=torch.int32)

    def _triangle_part(self, x, params, ts):
        """One part of a symmetric all-pairs sweep: every unordered pair {i, j} belongs to exactly one part
        (here: by the 64-row block of min(i, j), dealt out round-robin -- the kernel's own deal differs, any
        exact partition merges to the same lists) and feeds the lists of both i and j."""
        part, parts = params["part_index"], params["part_count"]
        n = x.shape[0]
        xn = O.l2_normalize(x)
        S = xn @ xn.T
        if ts is not None:
            S[O.time_excluded(ts, ts, params["gap"])] = -np.inf
        i, j = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
        S[((np.minimum(i, j) // 64) % parts) != part] = -np.inf
        S[S < np.float32(params["thr"])] = -np.inf
        sc, ix, ct = O._topk_rows(S, params["k"])
        self._mode = 2 if self.overflow_on_rank == part else 1
        return _Res(keys=torch.from_numpy(O.pack_keys(sc, ix)))

    def gated_topk(self, q_bf16, db_bf16, params, q_ts=None, db_ts=None, q_floor=None, db_floor=None,
                   want_keys=False, want_lists=True):
        if params.get("part_count", 0) > 1:
            assert q_bf16 is db_bf16 and q_ts is db_ts and q_floor is db_floor and params["symmetric"] == 1
            return self._triangle_part(q_bf16.numpy(), params, None if q_ts is None else q_ts.numpy())
        self._mode = 0
        lo, n = params["offset"], db_bf16.shape[0]
        r = O.gated_topk(q_bf16.numpy(), db_bf16.numpy(), None if q_ts is None else q_ts.numpy(),
                         None if db_ts is None else db_ts.numpy(), None if q_floor is None else q_floor.numpy(),
                         None if db_floor is None else db_floor.numpy(), k=params["k"], threshold=params["thr"],
                         min_time_gap=params["gap"], max_floor_diff=params["mfd"], db_index_offset=lo, normalize=True)
        keys = torch.from_numpy(O.pack_keys(r["scores"], r["idx"]))
        return _Res(keys=keys, scores=torch.from_numpy(r["scores"]), idx=torch.from_numpy(r["idx"]),
                    valid=torch.from_numpy(r["valid"]), count=torch.from_numpy(r["count"]))

    def merge_topk(self, keys_gathered, k, q_floor=None, db_floor_all=None, max_floor_diff=-1):
        merged = O.merge_keys(keys_gathered.numpy(), k)
        sc, ix = O.unpack_keys(merged)
        got = ix >= 0
        valid = got.copy()
        if q_floor is not None and db_floor_all is not None and max_floor_diff >= 0:
            mf = db_floor_all.numpy()[np.where(got, ix, 0)]
            valid = got & O.floor_ok(q_floor.numpy()[:, None], mf, max_floor_diff)
        return _Res(scores=torch.from_numpy(sc), idx=torch.from_numpy(ix), valid=torch.from_numpy(valid),
                    count=torch.from_numpy(got.sum(axis=1).astype(np.int32)), keys=torch.from_numpy(merged))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n_db, n_q, d, k = 1301, 257, 48, 9
        desc, ts, fl = synthetic.make_case(n_db, d, 4, seed=3)
        fl = fl.astype(np.int32)
        lo, hi = shard_bounds(n_db, world, rank)
        sr = ShardedRetrieval(OracleEngine(desc[:n_q], desc))
        assert sr.world == world and sr.rank == rank
        t = torch.from_numpy
        res = sr.sweep(t(desc[:n_q]), t(desc[lo:hi]), lambda off: dict(offset=off, k=k, thr=0.3, gap=5.0, mfd=0), lo,
                       q_ts=t(ts[:n_q]), db_ts_shard=t(ts[lo:hi]), q_floor=t(fl[:n_q]), db_floor_shard=t(fl[lo:hi]),
                       db_floor_all=t(fl), max_floor_diff=0)
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), scores=res.scores.numpy(), idx=res.idx.numpy(),
                 valid=res.valid.numpy(), count=res.count.numpy())
        # host-input form: only rank 0 holds the queries, the others receive them by broadcast
        mk = lambda off: dict(offset=off, k=k, thr=0.3, gap=5.0, mfd=0)
        res2 = sr.sweep_from_host(t(desc[:n_q].copy()) if rank == 0 else None, t(desc[lo:hi].copy()), t(ts), t(fl), mk, lo, hi,
                                  n_q, max_floor_diff=0)
        for a, b in ((res.scores, res2.scores), (res.idx, res2.idx), (res.valid, res2.valid), (res.count, res2.count)):
            assert torch.equal(a, b), "sweep_from_host differs from sweep"
        # all-pairs sweep: the ranks split the triangle of pairs; then the same with one rank reporting an
        # overflowed part, which must send every rank down the row-sharded path
        n_ap = 700
        x, tts, tfl = t(desc[:n_ap]), t(ts[:n_ap]), t(fl[:n_ap])
        mk_sym = lambda off: dict(offset=off, k=k, thr=0.3, gap=5.0, mfd=0, symmetric=1)   # small case: force the triangle split
        for knob, how in ((None, "triangle"), (1, "rows (candidate buffers overflowed)")):
            sr.engine.overflow_on_rank = knob
            ap = sr.sweep_all_pairs(x, mk_sym, ts=tts, floor=tfl, max_floor_diff=0)
            assert sr.last_all_pairs == how
            # deferred form: the same lists, the flag looked at when the result is asked for
            pend = sr.sweep_all_pairs(x, mk_sym, ts=tts, floor=tfl, max_floor_diff=0, defer=True)
            ap_d = pend.result()
            assert pend.overflowed == (knob is not None) and torch.equal(ap_d.idx, ap.idx) and torch.equal(ap_d.count, ap.count)
            # this rank's share of the rows only
            part = sr.sweep_all_pairs(x, mk_sym, ts=tts, floor=tfl, max_floor_diff=0, gather=False)
            plo, phi = shard_bounds(n_ap, world, rank)
            assert (part.lo, part.hi) == (plo, phi) and torch.equal(part.result.idx, ap.idx[plo:phi])
            assert torch.equal(part.result.valid, ap.valid[plo:phi]) and torch.equal(part.result.count, ap.count[plo:phi])
            np.savez(os.path.join(out_dir, f"ap{0 if knob is None else 1}_rank{rank}.npz"), scores=ap.scores.numpy(),
                     idx=ap.idx.numpy(), valid=ap.valid.numpy(), count=ap.count.numpy())
        sr.engine.overflow_on_rank = None
        alo, ahi = shard_bounds(n_ap, world, rank)
        ap2 = sr.sweep_all_pairs_from_host(t(desc[alo:ahi].copy()), tts, tfl, mk_sym, alo, ahi, n_ap, max_floor_diff=0)
        assert sr.last_all_pairs == "triangle"
        assert torch.equal(ap2.idx, ap.idx) and torch.equal(ap2.count, ap.count), "host-input all-pairs form differs"
        # the automatic rule leaves short sweeps / short descriptors to the row-sharded full sweep: same lists
        ap3 = sr.sweep_all_pairs(x, mk, ts=tts, floor=tfl, max_floor_diff=0)
        assert sr.last_all_pairs == "rows" and torch.equal(ap3.idx, ap.idx) and torch.equal(ap3.count, ap.count)
        # a rank's share of a rectangular sweep
        part = sr.sweep(t(desc[:n_q]), t(desc[lo:hi]), mk, lo, q_ts=t(ts[:n_q]), db_ts_shard=t(ts[lo:hi]), q_floor=t(fl[:n_q]),
                        db_floor_shard=t(fl[lo:hi]), db_floor_all=t(fl), max_floor_diff=0, gather=False)
        qlo, qhi = shard_bounds(n_q, world, rank)
        assert (part.lo, part.hi) == (qlo, qhi) and torch.equal(part.result.idx, res.idx[qlo:qhi])
    finally:
        dist.destroy_process_group()


def test_shard_bounds_cover_exactly():
    for n in (0, 1, 7, 1000, 1000003):
        for w in (1, 2, 3, 8):
            b = [shard_bounds(n, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def test_key_roundtrip_and_order():
    rng = np.random.default_rng(0)
    s = rng.standard_normal((50, 8)).astype(np.float32)
    s[0, :3] = [0.0, -0.0, np.inf]
    s[1, :2] = 0.25                                     # tie: lower index must win
    ix = rng.permutation(400).reshape(50, 8).astype(np.int64)
    ix[1, :2] = [7, 3]
    keys = O.pack_keys(s, ix)
    s2, i2 = O.unpack_keys(keys)
    assert np.array_equal(s2, s) and np.array_equal(i2, ix)
    ku = keys.view(np.uint64)
    assert ku[1, 1] > ku[1, 0]                          # same score, index 3 beats index 7
    order = np.argsort(-ku[2].astype(np.float64), kind="stable")
    assert np.all(np.diff(s[2][order]) <= 0)
    empty = O.pack_keys(np.array([-np.inf], np.float32), np.array([-1]))
    assert empty[0] == 0 and O.unpack_keys(empty)[1][0] == -1


def test_sharded_sweep_world2_gloo(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    n_db, n_q, d, k = 1301, 257, 48, 9
    desc, ts, fl = synthetic.make_case(n_db, d, 4, seed=3)
    fl = fl.astype(np.int32)
    whole = O.gated_topk(desc[:n_q], desc, ts[:n_q], ts, fl[:n_q], fl, k=k, threshold=0.3, min_time_gap=5.0,
                         max_floor_diff=0)
    r0 = np.load(tmp_path / "rank0.npz")
    r1 = np.load(tmp_path / "rank1.npz")
    for key in ("scores", "idx", "valid", "count"):
        assert np.array_equal(r0[key], r1[key]), f"ranks disagree on {key}"
    # shard sweeps normalise their own slice: identical arithmetic per row -> identical result
    assert np.array_equal(r0["idx"], whole["idx"])
    assert np.array_equal(r0["scores"], whole["scores"])
    assert np.array_equal(r0["valid"], whole["valid"])
    assert np.array_equal(r0["count"], whole["count"])
    # all-pairs over the first 700 keyframes, triangle split and its row-sharded fallback
    ap = O.gated_topk(desc[:700], desc[:700], ts[:700], ts[:700], fl[:700], fl[:700], k=k, threshold=0.3, min_time_gap=5.0,
                      max_floor_diff=0)
    for tag in ("ap0", "ap1"):
        for rank in (0, 1):
            r = np.load(tmp_path / f"{tag}_rank{rank}.npz")
            assert np.array_equal(r["idx"], ap["idx"]) and np.array_equal(r["valid"], ap["valid"]), (tag, rank)
            assert np.allclose(r["scores"], ap["scores"], atol=1e-6) and np.array_equal(r["count"], ap["count"])
