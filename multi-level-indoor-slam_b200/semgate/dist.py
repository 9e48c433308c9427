"""Row-sharded database across the GPUs of one box.

One process per GPU (`torch.distributed`, NCCL over NVLink/NVSwitch).  Every rank
holds all queries and a contiguous slice of database rows; it runs the fused
sweep on its slice with global column indices (`db_index_offset`), then the
per-rank `[Q,k]` candidate-key lists (8*Q*k bytes per rank) are exchanged and merged
by the K3 kernel on every rank — either all-gathered first (NCCL) or read in place
from the peers' memory over NVLink by the merge kernel itself.  Top-k under a total
order is an associative merge, so the result equals the single-GPU sweep exactly.

That exchange is the path's only one; there is no reduction over the descriptor
dimension and no all-to-all.

All-pairs sweeps (`sweep_all_pairs`: the queries are the whole database, held by every
rank) split the triangle of similarity tiles instead of the rows, because S = S^T:
half the tensor work, the same exchange.
"""
from __future__ import annotations

from typing import Optional, Tuple


def shard_bounds(n_total: int, world: int, rank: int) -> Tuple[int, int]:
    """Balanced contiguous split of database rows: rank r owns [lo, hi)."""
    if not (0 <= rank < world):
        raise ValueError("rank outside the group")
    return (n_total * rank) // world, (n_total * (rank + 1)) // world


class ShardedRetrieval:
    """Gated top-k over a database sharded by rows across the ranks of `group`.

    exchange = "allgather": the per-rank `[Q,k]` key lists are all-gathered (NCCL; gloo in the CPU
    tests) and merged.  exchange = "peer": every rank writes its keys into a symmetric-memory buffer
    that the other ranks have mapped, and after a cross-GPU barrier on the stream the merge kernel
    reads the G lists IN PLACE over NVLink (`semgate_merge_topk_peers`) — one kernel does the
    exchange and the merge, the gathered copy never exists.  "auto" = "peer" on NCCL groups when
    symmetric memory can be set up, else "allgather"; both give bit-identical results."""

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
        self._symm = None          # (shape, keys tensor, handle)
        self.peer_error = None     # why "auto" fell back, if it did
        self.last_all_pairs = None # how the last sweep_all_pairs ran

    # -- peer-memory exchange ----------------------------------------------------
    def _symm_keys(self, Q: int, k: int, device):
        """Symmetric `[Q,k]` int64 buffer + rendezvous handle, cached per shape."""
        import torch
        import torch.distributed._symmetric_memory as symm_mem
        if self._symm is None or self._symm[0] != (Q, k):
            t = symm_mem.empty((Q, k), dtype=torch.int64, device=device)
            grp = self.group if self.group is not None else self.dist.group.WORLD
            hdl = symm_mem.rendezvous(t, grp)
            self._symm = ((Q, k), t, hdl)
        return self._symm[1], self._symm[2]

    def _use_peer(self) -> bool:
        if self.exchange == "allgather" or self.world == 1:
            return False
        if self.exchange == "peer":
            return True
        if self.peer_error is not None:
            return False
        try:
            return self.dist.get_backend(self.group) == "nccl"
        except Exception:
            return False

    def _exchange_and_merge(self, local_fn, Q: int, k: int, device, q_floor, db_floor_all, max_floor_diff: int,
                            after_local=None):
        """`local_fn(keys)` runs this rank's sweep, writing its `[Q,k]` key lists into `keys` when given
        (the symmetric-memory buffer) or returning them in `.keys`; the lists of all ranks are then merged
        on every rank.  `after_local()`, if given, is called by every rank right after its sweep is queued
        (stream-ordered work only: nothing here waits for the GPU)."""
        import torch
        if self._use_peer():
            try:
                keys, hdl = self._symm_keys(Q, k, device)
            except Exception as e:          # no P2P mapping on this system
                if self.exchange == "peer":
                    raise
                self.peer_error = f"{type(e).__name__}: {e}"
            else:
                hdl.barrier(channel=0)      # every rank has finished reading the previous step's keys
                local_fn(keys)
                if after_local is not None:
                    after_local()
                hdl.barrier(channel=1)      # every rank's keys are written
                return self.engine.merge_topk_peers(hdl.buffer_ptrs_dev, self.world, Q, k, q_floor=q_floor,
                                                    db_floor_all=db_floor_all, max_floor_diff=max_floor_diff)
        local = local_fn(None)
        if after_local is not None:
            after_local()
        buf = self._gather_buf
        if buf is None or buf.shape != (self.world * Q, k) or buf.device != local.keys.device:
            # concatenation along dim 0 is the layout every backend accepts; viewed as [G,Q,k] below
            buf = torch.empty((self.world * Q, k), dtype=torch.int64, device=local.keys.device)
            self._gather_buf = buf
        self.dist.all_gather_into_tensor(buf, local.keys, group=self.group)
        return self.engine.merge_topk(buf.view(self.world, Q, k), k, q_floor=q_floor, db_floor_all=db_floor_all,
                                      max_floor_diff=max_floor_diff)

    def sweep(self, q_bf16, db_shard_bf16, make_params, shard_lo: int, q_ts=None, db_ts_shard=None, q_floor=None,
              db_floor_shard=None, db_floor_all=None, max_floor_diff: int = -1):
        """`make_params(db_index_offset)` builds the sweep parameters for this shard.
        Returns the merged TopkResult (identical on every rank)."""
        params = make_params(shard_lo)
        eng = self.engine
        if self.world == 1:
            return eng.gated_topk(q_bf16, db_shard_bf16, params, q_ts=q_ts, db_ts=db_ts_shard, q_floor=q_floor,
                                  db_floor=db_floor_shard)
        Q, k = q_bf16.shape[0], params.k if hasattr(params, "k") else params["k"]

        def local_fn(keys):
            if keys is not None:
                return eng.gated_topk(q_bf16, db_shard_bf16, params, q_ts=q_ts, db_ts=db_ts_shard, q_floor=q_floor,
                                      db_floor=db_floor_shard, want_lists=False, keys=keys)
            return eng.gated_topk(q_bf16, db_shard_bf16, params, q_ts=q_ts, db_ts=db_ts_shard, q_floor=q_floor,
                                  db_floor=db_floor_shard, want_keys=True, want_lists=False)
        return self._exchange_and_merge(local_fn, Q, k, q_bf16.device, q_floor, db_floor_all, max_floor_diff)

    def sweep_all_pairs(self, x_bf16, make_params, ts=None, floor=None, max_floor_diff: int = -1, compact: bool = False):
        """All-pairs sweep of a database every rank holds in full (`find_loop_closures` over the whole map:
        the queries ARE the database, place_recognition.py:190 computes X X^T).  Similarity is symmetric, so
        the ranks split the TRIANGLE of tiles instead of the rows: every rank computes its share of the
        tiles on or above the block diagonal once and gates each of them in both directions
        (`semgate_topk_params.part_index / part_count`), then the per-rank lists are merged as in `sweep`.
        Half the tensor work of the row-sharded sweep, same lists.  If any rank's candidate buffers overflow
        (thresholds that admit most of the database) all ranks agree on it and redo the sweep row-sharded.
        Returns the merged TopkResult (identical on every rank); with `compact=True` the flat candidate
        arrays of `engine.compact(result)` instead (queued before the host looks at the overflow flag, so the
        GPU never waits for the host)."""
        eng = self.engine
        n = x_bf16.shape[0]
        params = make_params(0)
        done = (lambda r: eng.compact(r)) if compact else (lambda r: r)
        if self.world == 1:
            return done(eng.gated_topk(x_bf16, x_bf16, params, q_ts=ts, db_ts=ts, q_floor=floor, db_floor=floor))
        k = params.k if hasattr(params, "k") else params["k"]
        wanted = (params.symmetric if hasattr(params, "symmetric") else params.get("symmetric", 0)) >= 0

        def rows():
            lo, hi = shard_bounds(n, self.world, self.rank)
            return self.sweep(x_bf16, x_bf16[lo:hi], make_params, lo, q_ts=ts, db_ts_shard=None if ts is None else ts[lo:hi],
                              q_floor=floor, db_floor_shard=None if floor is None else floor[lo:hi], db_floor_all=floor,
                              max_floor_diff=max_floor_diff)
        if not wanted or n <= 256:
            return done(rows())
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

        # Did any rank's candidate buffers overflow?  The flag is copied and all-reduced on the device while the
        # exchange and the merge are being queued behind it; the host reads it once, when everything is in flight.
        over = []

        def after_local():
            over.append(eng.last_sweep_overflow())
            self.dist.all_reduce(over[0], op=self.dist.ReduceOp.MAX, group=self.group)
        res = done(self._exchange_and_merge(local_fn, n, k, x_bf16.device, floor, floor, max_floor_diff, after_local=after_local))
        if int(over[0].item()) != 0:          # incomplete lists somewhere: every rank redoes the sweep row-sharded
            self.last_all_pairs = "rows (candidate buffers overflowed)"
            return done(rows())
        self.last_all_pairs = "triangle"
        return res

    def sweep_all_pairs_from_host(self, x_shard_host, ts_all_host, floor_all_host, make_params, shard_lo: int,
                                  shard_hi: int, n: int, max_floor_diff: int = -1, compact: bool = False):
        """End-to-end form of `sweep_all_pairs` for a database that lives in pinned HOST memory, one
        contiguous row shard per rank (equal shards): every rank uploads and normalises only its own rows
        (fp32 `[hi-lo, D]`), the normalised bf16 rows meet on every GPU through an NCCL all-gather over
        NVLink, then the ranks split the triangle of tiles.  Returns the merged TopkResult."""
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
        return self.sweep_all_pairs(buf, make_params, ts=ts, floor=fl, max_floor_diff=max_floor_diff, compact=compact)

    def sweep_from_host(self, q_host, db_shard_host, ts_all_host, floor_all_host, make_params, shard_lo: int,
                        shard_hi: int, n_q: int, max_floor_diff: int = -1, src: int = 0):
        """End-to-end form of `sweep` for inputs that live in pinned HOST memory.

        Every rank uploads and normalises only its own database shard (fp32 `[hi-lo, D]`); the query
        descriptors (fp32 `[n_q, D]`, given on rank `src` only, may be the same tensor as its shard)
        cross PCIe once, on `src`, and reach the other GPUs as normalised bf16 rows through an NCCL
        broadcast over NVLink instead of `world` more host copies.  Timestamps / floor labels (12 B per
        keyframe) are uploaded by every rank.  Returns the merged TopkResult (identical on every rank)."""
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
                          db_floor_all=fl, max_floor_diff=max_floor_diff)
