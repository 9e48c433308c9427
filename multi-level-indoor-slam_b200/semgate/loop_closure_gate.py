"""Host-side mirror of the reference's floor gate, backed by libsemgate.

Interface of scripts/semantic_gating/loop_closure_gate.py in the reference:
LoopClosureCandidate (:16-25) and SemanticLoopClosureGate (:28-148) with the same
constructor, methods, counter names and printed summary.  `gate_candidates` sends
the whole batch through one kernel (`semgate_gate_candidates_host`); the integer
decisions are identical to the reference's per-candidate Python loop.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Sequence, Tuple

import numpy as np

from . import _native


@dataclass
class LoopClosureCandidate:
    """A potential loop closure."""
    query_idx: int
    match_idx: int
    similarity_score: float
    query_floor: int
    match_floor: int
    is_valid: bool = True
    rejection_reason: str = ""


class SemanticLoopClosureGate:
    """Reject loop-closure candidates whose keyframes lie on different floors."""

    def __init__(self, floor_labels: np.ndarray, strict_mode: bool = True, device: int = 0):
        """
        floor_labels: floor label of every keyframe.
        strict_mode:  True rejects any cross-floor pair; False only pairs whose floors
                      differ by more than one (reference :42-50).
        """
        self.floor_labels = floor_labels
        self.strict_mode = strict_mode
        self.device = device
        self.stats = {'total_candidates': 0, 'accepted': 0, 'rejected_cross_floor': 0, 'rejected_other': 0}

    # -- helpers ---------------------------------------------------------------
    def _max_diff(self) -> int:
        return 0 if self.strict_mode else 1

    def _labels_i32(self) -> np.ndarray:
        fl = np.asarray(self.floor_labels)
        if fl.size and (fl.max() > 2**31 - 1 or fl.min() < -2**31 + 1):
            raise ValueError("floor labels must fit in int32")
        return np.ascontiguousarray(fl, dtype=np.int32)

    def _reason(self, qf, mf) -> str:
        return f"Cross-floor: {qf} vs {mf}" if self.strict_mode else f"Floor diff > 1: {qf} vs {mf}"

    def gate_arrays(self, query_idx: Sequence[int], match_idx: Sequence[int]) -> np.ndarray:
        """Vector form: bool[M] acceptance mask; updates `stats` like M calls of gate_candidate."""
        qi = np.ascontiguousarray(query_idx, dtype=np.int64)
        mi = np.ascontiguousarray(match_idx, dtype=np.int64)
        n = len(self.floor_labels)
        # Python negative indexing is legal in the reference (numpy wraps); normalise before the kernel
        qi = np.where(qi < 0, qi + n, qi)
        mi = np.where(mi < 0, mi + n, mi)
        if qi.size and (qi.min() < 0 or mi.min() < 0 or qi.max() >= n or mi.max() >= n):
            raise IndexError("candidate index out of bounds for floor_labels")   # numpy raises IndexError too
        eng = _native.get_engine(self.device)
        ok, accepted, rejected = eng.gate_candidates_host(self._labels_i32(), qi.astype(np.int32), mi.astype(np.int32),
                                                          self._max_diff())
        self.stats['total_candidates'] += int(qi.size)
        self.stats['accepted'] += accepted
        self.stats['rejected_cross_floor'] += rejected
        return ok

    # -- reference interface -----------------------------------------------------
    def gate_candidate(self, query_idx: int, match_idx: int, similarity_score: float = 0.0) -> LoopClosureCandidate:
        ok = bool(self.gate_arrays([query_idx], [match_idx])[0])
        qf, mf = self.floor_labels[query_idx], self.floor_labels[match_idx]
        return LoopClosureCandidate(query_idx, match_idx, similarity_score, qf, mf, ok, "" if ok else self._reason(qf, mf))

    def gate_candidates(self, candidates: List[Tuple[int, int, float]]) -> Tuple[List, List]:
        """(valid, rejected) lists of LoopClosureCandidate, input order preserved (reference :105-126)."""
        if len(candidates) == 0:
            return [], []
        qi = [c[0] for c in candidates]
        mi = [c[1] for c in candidates]
        ok = self.gate_arrays(qi, mi)
        fl = self.floor_labels
        valid, rejected = [], []
        for (q, m, score), good in zip(candidates, ok.tolist()):
            qf, mf = fl[q], fl[m]
            if good:
                valid.append(LoopClosureCandidate(q, m, score, qf, mf, True, ""))
            else:
                rejected.append(LoopClosureCandidate(q, m, score, qf, mf, False, self._reason(qf, mf)))
        return valid, rejected

    def get_stats(self) -> Dict:
        total = self.stats['total_candidates']
        if total > 0:
            self.stats['acceptance_rate'] = self.stats['accepted'] / total
            self.stats['rejection_rate'] = 1 - self.stats['acceptance_rate']
        return self.stats

    def print_summary(self):
        stats = self.get_stats()
        print("\n" + "=" * 50)
        print("LOOP CLOSURE GATING SUMMARY")
        print("=" * 50)
        print(f"Total candidates:      {stats['total_candidates']}")
        print(f"Accepted:              {stats['accepted']}")
        print(f"Rejected (cross-floor): {stats['rejected_cross_floor']}")
        if stats['total_candidates'] > 0:
            print(f"Acceptance rate:       {stats['acceptance_rate']:.1%}")
            print(f"Perceptual aliasing prevented: {stats['rejected_cross_floor']}")
        print("=" * 50)
