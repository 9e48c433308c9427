"""`torch.ops.semgate.*` — the path as PyTorch operators.

A thin C++ op library (`csrc/torch_ops.cpp` -> `libsemgate_torch.so`, built in-tree by
`build.py --torch-ops` / `__graft_entry__.build()`) registers four operators over libsemgate's C ABI:

    normalize_cast(x f32[N,D]) -> bf16[N,pad64(D)]
    gated_topk(q_bf16, db_bf16, q_floor?, db_floor?, q_ts?, db_ts?, min_time_gap, threshold, k,
               max_floor_diff, gate_mode, db_index_offset) -> (scores, idx, valid, count, keys)
    merge_topk(keys[G,Q,k], q_floor?, db_floor_all?, max_floor_diff) -> (scores, idx, valid, count, keys)
    compact(scores, idx, valid, count) -> (query_idx, match_idx, similarity, is_valid, total[1])

They allocate with torch's caching allocator and launch on torch's current CUDA stream; CUDA tensors
only (there is no CPU implementation to dispatch to).  `semgate._native.Engine` (ctypes) is the same
path without the op registry; both end in the same kernels.
"""
from __future__ import annotations

import os

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libsemgate_torch.so")
_ops = None


def load():
    """Register the operators (once) and return `torch.ops.semgate`."""
    global _ops
    if _ops is None:
        import torch
        if not os.path.isfile(LIB_PATH):
            raise ImportError(f"{LIB_PATH} not found. Build it with "
                              "`python multi-level-indoor-slam_b200/build.py --torch-ops`.")
        torch.ops.load_library(LIB_PATH)
        _ops = torch.ops.semgate
    return _ops
