"""Spatial-proximity loop-closure candidates + floor gating on pose arrays.

The compute steps of the reference's per-algorithm integration scripts
(`detect_loop_closure_candidates`, orb_slam3_integration.py:167-217, and
`apply_floor_gating`, :219-281; identical twins in droid_slam_integration.py and
lego_loam_integration.py), without their file loading, printing and plotting.
Both run on the GPU through libsemgate (radius-join kernel + gate kernel).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Tuple

import numpy as np

from . import _native
from .loop_closure_gate import SemanticLoopClosureGate


def detect_loop_closure_candidates_arrays(positions: np.ndarray, distance_threshold: float = 2.0,
                                          min_time_gap: int = 100, device: int = 0):
    """(query_idx, match_idx, distance) arrays: pairs i < j with ||p_i - p_j|| <= distance_threshold
    (fp64) and |i - j| >= min_time_gap, sorted by (i, j)."""
    eng = _native.get_engine(device)
    return eng.spatial_candidates_host(np.asarray(positions, dtype=np.float64)[:, :3], distance_threshold, min_time_gap)


def detect_loop_closure_candidates(positions: np.ndarray, distance_threshold: float = 2.0,
                                   min_time_gap: int = 100, device: int = 0) -> List[Tuple[int, int, float]]:
    """Same as a list of `(query_idx, match_idx, distance)` tuples (the reference's return type)."""
    i, j, d = detect_loop_closure_candidates_arrays(positions, distance_threshold, min_time_gap, device)
    return list(zip(i.tolist(), j.tolist(), d.tolist()))


@dataclass
class LoopClosureAnalysis:
    """Counters of apply_floor_gating (reference: orb_slam3_integration.py:30-37)."""
    total_candidates: int = 0
    same_floor_candidates: int = 0
    cross_floor_candidates: int = 0
    cross_floor_pairs: List[Tuple[int, int, int, int]] = field(default_factory=list)


def apply_floor_gating(query_idx, match_idx, floor_labels: np.ndarray, strict_mode: bool = True, device: int = 0,
                       max_pairs: int = 0):
    """Gate candidate pairs by floor.  Returns (LoopClosureAnalysis, gate, is_valid[M]).
    `same_floor`/`cross_floor` count label equality (reference :241-251); the gate's own
    counters follow strict / non-strict mode.  `max_pairs` > 0 also lists that many
    cross-floor pairs as (i, j, floor_i, floor_j)."""
    qi = np.asarray(query_idx, dtype=np.int64)
    mi = np.asarray(match_idx, dtype=np.int64)
    gate = SemanticLoopClosureGate(floor_labels, strict_mode=strict_mode, device=device)
    ok = gate.gate_arrays(qi, mi)
    if strict_mode:
        same = ok
    else:
        same = SemanticLoopClosureGate(floor_labels, strict_mode=True, device=device).gate_arrays(qi, mi)
    a = LoopClosureAnalysis(total_candidates=int(qi.size), same_floor_candidates=int(same.sum()),
                            cross_floor_candidates=int(qi.size - same.sum()))
    if max_pairs > 0:
        fl = np.asarray(floor_labels)
        cross = np.nonzero(~same)[0][:max_pairs]
        a.cross_floor_pairs = [(int(qi[c]), int(mi[c]), int(fl[qi[c]]), int(fl[mi[c]])) for c in cross]
    return a, gate, ok
