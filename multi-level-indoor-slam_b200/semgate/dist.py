"""Database sharded across the GPUs of one box.

One process per GPU (`torch.distributed`, NCCL over NVLink/NVSwitch).  Two partitions of the work:

* rectangular sweeps (queries != database): every rank holds all queries and a contiguous slice of
  database rows; it runs the fused sweep on its slice with global column indices (`db_index_offset`);
* all-pairs sweeps (`sweep_all_pairs`: the queries ARE the database, held by every rank): the ranks split
  the TRIANGLE of similarity tiles instead of the rows, because S = S^T — half the tensor work.

Either way every rank ends up with a `[Q,k]` list of candidate keys per query over ITS share of the pairs,
and the path's only exchange step is the merge of those G lists per query (top-k under a total order is an
associative merge, so the result equals the single-GPU sweep exactly).  There is no reduction over the
descriptor dimension and no all-to-all.

Exchange, `exchange="peer"` (default on NCCL groups): every rank's sweep writes its keys straight into a
symmetric-memory buffer the other ranks have mapped; ONE cross-GPU barrier on the stream orders the writes; then
every rank merges ITS OWN ROWS `[r*Q/G, (r+1)*Q/G)` of the G lists, reading them in place over NVLink
(`semgate_merge_topk_peers_rows`) — one kernel does the exchange and the merge, each key crosses NVLink once
(the replicated merge of round 1 read every key on every rank), and the same kernel ORs the ranks' overflow
flags (a word behind each rank's keys), so no collective is spent on them.  The key buffers are double-buffered:
step t writes buffer t%2, and a rank can only pass barrier t after every peer has queued its merge of step t-1,
so one barrier per step is enough.  Nothing in a step waits for the host: the overflow flag of a split
symmetric sweep is copied to pinned host memory behind the merge and looked at when the caller asks for the
result (`defer=True`: whenever the caller likes, e.g. one step late).

`exchange="allgather"`: NCCL (gloo in the CPU tests) all-gather of the key lists and a replicated merge; also
what "auto" falls back to, on all ranks together, when symmetric memory cannot be set up.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple


def shard_bounds(n_total: int, world: int, rank: int) -> Tuple[int, int]:
    """Balanced contiguous split of rows: rank r owns [lo, hi)."""
    if not (0 <= rank < world):
        raise ValueError("rank outside the group")
    return (n_total * rank) // world, (n_total * (rank + 1)) // world


@dataclass
class RowsResult:
    """A rank's share of a sharded sweep: the merged lists of query rows [lo, hi) (`gather=False`)."""
    lo: int
    hi: int
    result: "object"       # TopkResult with [hi-lo, k] arrays; idx are global database indices


class PendingSweep:
    """Result of `sweep_all_pairs(..., defer=True)`: everything is queued on the stream, the overflow flag of
    the split symmetric sweep is on its way to pinned host memory.  `result()` waits for the flag only and, in
    the rare case that some rank's candidate buffers overflowed, redoes the sweep row-sharded (all ranks take
    the same decision: every rank sees the OR of all flags)."""

    def __init__(self, value, flag_host, event, redo, owner):
        self._value, self._flag, self._event, self._redo, self._owner = value, flag_host, event, redo, owner
        self.overflowed = None

    def forget_value(self):
        """Drop the queued result (its device memory goes back to the allocator at once) but keep the flag: a caller
        that pipelines many sweeps and only wants to know afterwards whether any overflowed (`check()`)."""
        self._value = None
        return self

    def check(self) -> bool:
        """True iff this sweep's candidate buffers overflowed somewhere (waits for its flag only; no redo)."""
        if self._redo is None and self.overflowed is None:
            return False
        if self.overflowed is None:
            if self._event is not None:
                self._event.synchronize()
            self.overflowed = bool(int(self._flag[0]) != 0)
        return self.overflowed

    def result(self):
        if self._redo is not None:
            if self._event is not None:
                self._event.synchronize()
            self.overflowed = bool(int(self._flag[0]) != 0)
            if self.overflowed:
                self._owner.last_all_pairs = "rows (candidate buffers overflowed)"
                self._value = self._redo()
            self._redo = None
        return self._value


class ShardedRetrieval:
    """Gated top-k over a database sharded across the ranks of `group` (see the module docstring)."""

    def __init__(self, engine, group=None, exchange: str = "auto"):
        import torch.distributed as dist
        if exchange not in ("auto", "allgather", "peer"):
            raise ValueError("exchange must be 'auto', 'allgather' or 'peer'")
        self.engine = engine
        self.group = group
        self.dist = dist
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self._gather_buf = None
        self.exchange = exchange
        self._symm = None          # (shape, [(keys tensor, handle)] * 2)
        self._step = 0
        self._peer_ok = None       # agreed by all ranks at first use
        self.peer_error = None     # why "auto" fell back, if it did
        self.last_all_pairs = None # how the last sweep_all_pairs ran
        self._flag_host = None     # pinned int32 words for deferred overflow checks (one block, slots recycled)
        self._flag_next = 0

    # -- peer-memory exchange ----------------------------------------------------
    def _symm_keys(self, Q: int, k: int, device):
        """Two symmetric int64 buffers of Q*k keys + 8 trailing words (flag area), with their rendezvous
        handles, cached per shape."""
        import torch
        import torch.distributed._symmetric_memory as symm_mem
        if self._symm is None or self._symm[0] != (Q, k):
            grp = self.group if self.group is not None else self.dist.group.WORLD
            pairs = []
            for _ in range(2):
                t = symm_mem.empty((Q * k + 8,), dtype=torch.int64, device=device)
                t.zero_()
                pairs.append((t, symm_mem.rendezvous(t, grp)))
            self._symm = ((Q, k), pairs)
            self._step = 0
        return self._symm[1]

    def _use_peer(self, Q: int, k: int, device) -> bool:
        """Peer-memory exchange or all-gather: decided ONCE, by all ranks together (a rank that cannot map its
        peers must not leave the others waiting in a barrier)."""
        if self.exchange == "allgather" or self.world == 1:
            return False
        if self._peer_ok is None:
            import torch
            ok, err = 1, None
            try:
                if self.dist.get_backend(self.group) != "nccl":
                    ok = 0
            except Exception:
                ok = 0
            if ok:
                try:
                    self._symm_keys(Q, k, device)
                except Exception as e:          # no P2P mapping on this system
                    ok, err = 0, f"{type(e).__name__}: {e}"
            t = torch.tensor([ok], dtype=torch.int32, device=device)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN, group=self.group)
            self._peer_ok = bool(int(t.item()) == 1)
            if not self._peer_ok:
                self.peer_error = err or "a peer could not set up symmetric memory (or the backend is not NCCL)"
                self._symm = None
                if self.exchange == "peer":
                    raise RuntimeError(f"exchange='peer' is not available: {self.peer_error}")
        return self._peer_ok

    def _exchange_and_merge(self, local_fn, Q: int, k: int, device, q_floor, db_floor_all, max_floor_diff: int,
                            gather: bool, want_flag: bool):
        """`local_fn(keys)` runs this rank's sweep, writing its `[Q,k]` key lists into `keys` when given (the
        symmetric-memory buffer) or returning them in `.keys`.  Returns (result, flag): result = the merged lists
        of all rows (`gather`) or a RowsResult with this rank's rows; flag = int32 [1] device tensor holding the
        OR of all ranks' overflow flags (`want_flag`), else None.  Stream-ordered work only."""
        import torch
        eng = self.engine
        lo, hi = shard_bounds(Q, self.world, self.rank)
        if self._use_peer(Q, k, device):
            pairs = self._symm_keys(Q, k, device)
            buf, hdl = pairs[self._step & 1]
            self._step += 1
            local_fn(buf[:Q * k].view(Q, k))
            any_flag = None
            if want_flag:
                eng.last_sweep_overflow(out=buf[Q * k:Q * k + 1].view(torch.int32)[:1])
                any_flag = torch.empty((1,), dtype=torch.int32, device=device)
            hdl.barrier(channel=0)              # every rank's keys (and flag) are written
            mine = eng.merge_topk_peers_rows(hdl.buffer_ptrs_dev, self.world, Q, k, lo, hi - lo, q_floor=q_floor,
                                             db_floor_all=db_floor_all, max_floor_diff=max_floor_diff, want_keys=gather,
                                             flag_offset=Q * k, any_flag=any_flag)
            if not gather:
                return RowsResult(lo, hi, mine), any_flag
            return self._gather_rows(mine.keys, Q, k, q_floor, db_floor_all, max_floor_diff), any_flag
        local = local_fn(None)
        any_flag = None
        if want_flag:
            any_flag = eng.last_sweep_overflow()
            self.dist.all_reduce(any_flag, op=self.dist.ReduceOp.MAX, group=self.group)
        gbuf = self._gather_buf
        if gbuf is None or gbuf.shape != (self.world * Q, k) or gbuf.device != local.keys.device:
            # concatenation along dim 0 is the layout every backend accepts; viewed as [G,Q,k] below
            gbuf = torch.empty((self.world * Q, k), dtype=torch.int64, device=local.keys.device)
            self._gather_buf = gbuf
        self.dist.all_gather_into_tensor(gbuf, local.keys, group=self.group)
        res = eng.merge_topk(gbuf.view(self.world, Q, k), k, q_floor=q_floor, db_floor_all=db_floor_all,
                             max_floor_diff=max_floor_diff)
        if gather:
            return res, any_flag
        from ._native import TopkResult
        part = TopkResult(res.scores[lo:hi], res.idx[lo:hi], res.valid[lo:hi], res.count[lo:hi],
                          None if getattr(res, "keys", None) is None else res.keys[lo:hi])
        return RowsResult(lo, hi, part), any_flag

    def _gather_rows(self, my_keys, Q: int, k: int, q_floor, db_floor_all, max_floor_diff: int):
        """All ranks' merged row slices -> the full `[Q,k]` lists on every rank (NCCL all-gather of the merged
        keys, padded to the largest slice, then one decode pass)."""
        import torch
        per = -(-Q // self.world)
        send = my_keys
        if my_keys.shape[0] != per:
            send = torch.zeros((per, k), dtype=torch.int64, device=my_keys.device)
            send[:my_keys.shape[0]] = my_keys
        allk = torch.empty((self.world * per, k), dtype=torch.int64, device=my_keys.device)
        self.dist.all_gather_into_tensor(allk, send, group=self.group)
        if per * self.world != Q:
            parts = []
            for r in range(self.world):
                lo, hi = shard_bounds(Q, self.world, r)
                parts.append(allk[r * per:r * per + (hi - lo)])
            allk = torch.cat(parts, dim=0)
        return self.engine.merge_topk(allk.view(1, Q, k), k, q_floor=q_floor, db_floor_all=db_floor_all,
                                      max_floor_diff=max_floor_diff)

    @staticmethod
    def _param(params, name, default=0):
        return getattr(params, name) if hasattr(params, name) else params.get(name, default)

    def sweep(self, q_bf16, db_shard_bf16, make_params, shard_lo: int, q_ts=None, db_ts_shard=None, q_floor=None,
              db_floor_shard=None, db_floor_all=None, max_floor_diff: int = -1, gather: bool = True):
        """`make_params(db_index_offset)` builds the sweep parameters for this shard.
        Returns the merged TopkResult (identical on every rank), or with `gather=False` a RowsResult holding
        the merged lists of this rank's share of the query rows (no second exchange)."""
        params = make_params(shard_lo)
        eng = self.engine
        if self.world == 1:
            res = eng.gated_topk(q_bf16, db_shard_bf16, params, q_ts=q_ts, db_ts=db_ts_shard, q_floor=q_floor,
                                 db_floor=db_floor_shard)
            return res if gather else RowsResult(0, q_bf16.shape[0], res)
        Q, k = q_bf16.shape[0], self._param(params, "k")
        if k > 64:
            raise ValueError("sharded sweeps merge lists of at most 64 candidates per query")

        def local_fn(keys):
            if keys is not None:
                return eng.gated_topk(q_bf16, db_shard_bf16, params, q_ts=q_ts, db_ts=db_ts_shard, q_floor=q_floor,
                                      db_floor=db_floor_shard, want_lists=False, keys=keys)
            return eng.gated_topk(q_bf16, db_shard_bf16, params, q_ts=q_ts, db_ts=db_ts_shard, q_floor=q_floor,
                                  db_floor=db_floor_shard, want_keys=True, want_lists=False)
        return self._exchange_and_merge(local_fn, Q, k, q_bf16.device, q_floor, db_floor_all, max_floor_diff, gather, False)[0]

    def _finish(self, res, compact: bool):
        if not compact:
            return res
        if isinstance(res, RowsResult):
            return self.engine.compact(res.result, query_offset=res.lo)
        return self.engine.compact(res)

    def sweep_all_pairs(self, x_bf16, make_params, ts=None, floor=None, max_floor_diff: int = -1, compact: bool = False,
                        gather: bool = True, defer: bool = False):
        """All-pairs sweep of a database every rank holds in full (`find_loop_closures` over the whole map:
        the queries ARE the database, place_recognition.py:190 computes X X^T).  Similarity is symmetric, so
        the ranks split the TRIANGLE of tiles instead of the rows: every rank computes its share of the
        tiles on or above the block diagonal once and gates each of them in both directions
        (`semgate_topk_params.part_index / part_count`), then the per-rank lists are merged as in `sweep`.
        Half the tensor work of the row-sharded sweep, same lists; taken when the sweep is long enough to be
        tensor-bound (the library's own rule: n >= 8192 and 1024-d or longer; `symmetric = 1 / -1` in the
        parameters forces / forbids it), else the rows are sharded.  If any rank's candidate buffers overflow
        (thresholds that admit most of the database) all ranks learn it from the merge kernel and redo the sweep
        row-sharded.
        Returns the merged TopkResult (identical on every rank); `gather=False`: a RowsResult with this rank's
        share of the rows; `compact=True`: the flat candidate arrays of `engine.compact` of either (queued
        before the host looks at the overflow flag, so the GPU never waits for the host).
        `defer=True`: returns a PendingSweep at once; its `result()` does the one host read (a pipelined caller
        asks one step late and never stalls the GPU)."""
        eng = self.engine
        n = x_bf16.shape[0]
        params = make_params(0)
        if self.world == 1:
            res = eng.gated_topk(x_bf16, x_bf16, params, q_ts=ts, db_ts=ts, q_floor=floor, db_floor=floor)
            out = self._finish(res if gather else RowsResult(0, n, res), compact)
            return PendingSweep(out, None, None, None, self) if defer else out
        k = self._param(params, "k")
        sym = self._param(params, "symmetric")
        wanted = sym == 1 or (sym == 0 and n >= 8192 and x_bf16.shape[1] >= 1024)

        def make_rows_params(off):
            p = make_params(off)               # a row shard is not an all-pairs sweep: never symmetric
            if hasattr(p, "symmetric"):
                p.symmetric = -1
            else:
                p["symmetric"] = -1
            return p

        def rows():
            lo, hi = shard_bounds(n, self.world, self.rank)
            r = self.sweep(x_bf16, x_bf16[lo:hi], make_rows_params, lo, q_ts=ts, db_ts_shard=None if ts is None else ts[lo:hi],
                           q_floor=floor, db_floor_shard=None if floor is None else floor[lo:hi], db_floor_all=floor,
                           max_floor_diff=max_floor_diff, gather=gather)
            return self._finish(r, compact)
        if not wanted or n <= 256:
            self.last_all_pairs = "rows"
            out = rows()
            return PendingSweep(out, None, None, None, self) if defer else out
        if hasattr(params, "part_count"):
            params.symmetric, params.part_index, params.part_count = 1, self.rank, self.world
            if params.cta_group == 0:
                params.cta_group = 2          # a query block must be one database tile
        else:
            params.update(symmetric=1, part_index=self.rank, part_count=self.world)

        def local_fn(keys):
            if keys is not None:
                return eng.gated_topk(x_bf16, x_bf16, params, q_ts=ts, db_ts=ts, q_floor=floor, db_floor=floor,
                                      want_lists=False, keys=keys)
            return eng.gated_topk(x_bf16, x_bf16, params, q_ts=ts, db_ts=ts, q_floor=floor, db_floor=floor,
                                  want_keys=True, want_lists=False)

        res, any_flag = self._exchange_and_merge(local_fn, n, k, x_bf16.device, floor, floor, max_floor_diff, gather, True)
        out = self._finish(res, compact)
        flag_host, event = self._flag_to_host(any_flag)
        self.last_all_pairs = "triangle"
        pending = PendingSweep(out, flag_host, event, rows, self)
        return pending if defer else pending.result()

    def _flag_to_host(self, any_flag):
        """Copy of the agreed overflow flag into pinned host memory, queued behind the merge."""
        import torch
        if any_flag.device.type != "cuda":
            return any_flag, None
        # one pinned block for all flags, slots handed out round-robin (allocating pinned memory per step would cost
        # a millisecond and synchronise); 4096 sweeps may be pending before a slot comes round again
        if self._flag_host is None:
            self._flag_host = torch.zeros((4096,), dtype=torch.int32, pin_memory=True)
            self._flag_next = 0
        i = self._flag_next
        self._flag_next = (i + 1) % self._flag_host.shape[0]
        host = self._flag_host[i:i + 1]
        host.copy_(any_flag, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        return host, ev

    def sweep_all_pairs_from_host(self, x_shard_host, ts_all_host, floor_all_host, make_params, shard_lo: int,
                                  shard_hi: int, n: int, max_floor_diff: int = -1, compact: bool = False,
                                  gather: bool = True, defer: bool = False):
        """End-to-end form of `sweep_all_pairs` for a database that lives in pinned HOST memory, one
        contiguous row shard per rank (equal shards): every rank uploads and normalises only its own rows
        (fp32 `[hi-lo, D]`), the normalised bf16 rows meet on every GPU through an NCCL all-gather over
        NVLink, then the ranks split the triangle of tiles."""
        import torch
        eng = self.engine
        dev = getattr(eng, "torch_device", None) or torch.device("cuda", eng.device)
        if (shard_hi - shard_lo) * self.world != n:
            raise ValueError("sweep_all_pairs_from_host: the ranks' row shards must be equal")
        mine = eng.normalize_cast(x_shard_host.to(dev, non_blocking=True))
        if self.world > 1:
            buf = getattr(self, "_rows_buf", None)
            if buf is None or buf.shape != (n, mine.shape[1]) or buf.dtype != mine.dtype or buf.device != mine.device:
                buf = torch.empty((n, mine.shape[1]), dtype=mine.dtype, device=mine.device)
                self._rows_buf = buf
            self.dist.all_gather_into_tensor(buf, mine, group=self.group)
        else:
            buf = mine
        ts = ts_all_host.to(dev, non_blocking=True) if ts_all_host is not None else None
        fl = floor_all_host.to(dev, non_blocking=True) if floor_all_host is not None else None
        return self.sweep_all_pairs(buf, make_params, ts=ts, floor=fl, max_floor_diff=max_floor_diff, compact=compact,
                                    gather=gather, defer=defer)

    def sweep_from_host(self, q_host, db_shard_host, ts_all_host, floor_all_host, make_params, shard_lo: int,
                        shard_hi: int, n_q: int, max_floor_diff: int = -1, src: int = 0, gather: bool = True):
        """End-to-end form of `sweep` for inputs that live in pinned HOST memory.

        Every rank uploads and normalises only its own database shard (fp32 `[hi-lo, D]`); the query
        descriptors (fp32 `[n_q, D]`, given on rank `src` only, may be the same tensor as its shard)
        cross PCIe once, on `src`, and reach the other GPUs as normalised bf16 rows through an NCCL
        broadcast over NVLink instead of `world` more host copies.  Timestamps / floor labels (12 B per
        keyframe) are uploaded by every rank."""
        import torch
        eng = self.engine
        dev = getattr(eng, "torch_device", None) or torch.device("cuda", eng.device)
        same = self.rank == src and q_host is db_shard_host
        db_bf16 = eng.normalize_cast(db_shard_host.to(dev, non_blocking=True))
        if self.rank == src:
            q_bf16 = db_bf16 if same else eng.normalize_cast(q_host.to(dev, non_blocking=True))
            if q_bf16.shape[0] != n_q:
                q_bf16 = q_bf16[:n_q].contiguous()
        else:
            q_bf16 = torch.empty((n_q, db_bf16.shape[1]), dtype=db_bf16.dtype, device=dev)
        if self.world > 1:
            self.dist.broadcast(q_bf16, src=src, group=self.group)
        ts = ts_all_host.to(dev, non_blocking=True) if ts_all_host is not None else None
        fl = floor_all_host.to(dev, non_blocking=True) if floor_all_host is not None else None
        return self.sweep(q_bf16, db_bf16, make_params, shard_lo,
                          q_ts=None if ts is None else ts[:n_q], db_ts_shard=None if ts is None else ts[shard_lo:shard_hi],
                          q_floor=None if fl is None else fl[:n_q], db_floor_shard=None if fl is None else fl[shard_lo:shard_hi],
                          db_floor_all=fl, max_floor_diff=max_floor_diff, gather=gather)
