"""Row-sharded database across the GPUs of one box.

One process per GPU (`torch.distributed`, NCCL over NVLink/NVSwitch).  Every rank
holds all queries and a contiguous slice of database rows; it runs the fused
sweep on its slice with global column indices (`db_index_offset`), then the
per-rank `[Q,k]` candidate-key lists are all-gathered (8*Q*k bytes per rank) and
merged by the K3 kernel on every rank.  Top-k under a total order is an
associative merge, so the result equals the single-GPU sweep exactly.

The path's only exchange step is that all-gather; there is no reduction over the
descriptor dimension and no all-to-all.
"""
from __future__ import annotations

from typing import Optional, Tuple


def shard_bounds(n_total: int, world: int, rank: int) -> Tuple[int, int]:
    """Balanced contiguous split of database rows: rank r owns [lo, hi)."""
    if not (0 <= rank < world):
        raise ValueError("rank outside the group")
    return (n_total * rank) // world, (n_total * (rank + 1)) // world


class ShardedRetrieval:
    """Gated top-k over a database sharded by rows across the ranks of `group`."""

    def __init__(self, engine, group=None):
        import torch.distributed as dist
        self.engine = engine
        self.group = group
        self.dist = dist
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self._gather_buf = None

    def sweep(self, q_bf16, db_shard_bf16, make_params, shard_lo: int, q_ts=None, db_ts_shard=None, q_floor=None,
              db_floor_shard=None, db_floor_all=None, max_floor_diff: int = -1):
        """`make_params(db_index_offset)` builds the sweep parameters for this shard.
        Returns the merged TopkResult (identical on every rank)."""
        import torch
        params = make_params(shard_lo)
        if self.world == 1:
            return self.engine.gated_topk(q_bf16, db_shard_bf16, params, q_ts=q_ts, db_ts=db_ts_shard, q_floor=q_floor,
                                          db_floor=db_floor_shard)
        local = self.engine.gated_topk(q_bf16, db_shard_bf16, params, q_ts=q_ts, db_ts=db_ts_shard, q_floor=q_floor,
                                       db_floor=db_floor_shard, want_keys=True, want_lists=False)
        Q, k = local.keys.shape
        buf = self._gather_buf
        if buf is None or buf.shape != (self.world * Q, k) or buf.device != local.keys.device:
            # concatenation along dim 0 is the layout every backend accepts; viewed as [G,Q,k] below
            buf = torch.empty((self.world * Q, k), dtype=torch.int64, device=local.keys.device)
            self._gather_buf = buf
        self.dist.all_gather_into_tensor(buf, local.keys, group=self.group)
        return self.engine.merge_topk(buf.view(self.world, Q, k), k, q_floor=q_floor, db_floor_all=db_floor_all,
                                      max_floor_diff=max_floor_diff)
