"""CPU tier: a small executable model of the symmetric sweep's bookkeeping (gated_topk.cuh, SYM = true).

The kernel computes every similarity tile on or above the block diagonal once and reads it in two directions:
rows -> the owning run's k-list (admission bound = threshold until the list is full, then its k-th key), columns ->
candidates appended to the column keyframe's buffer if they pass that keyframe's *published* bound, which may be
stale or missing.  K3 then takes the top-k of a keyframe's lists and buffer.  The claim the kernel relies on: for
ANY tile order, ANY split of a block's tiles into runs and ANY staleness of the bounds, the result equals the
brute-force top-k of the full matrix under the total order (score descending, index ascending), ties included.
This model restates that bookkeeping in numpy/python and checks the claim on random instances; the kernel itself is
checked against the full sweep on the GPU tier (tests/test_gpu_parity.py::test_symmetric_sweep_*).
"""
import numpy as np
import pytest

from oracle import semgate_oracle as O


def _key(score, idx):
    """Total order of the kernel's packed keys: larger = better (score descending, index ascending)."""
    return (float(score), -int(idx))


class RunList:
    """RowList of common.cuh: k unsorted keys, admission bound = threshold until full, then the k-th score."""

    def __init__(self, k, thr):
        self.k, self.keys, self.f = k, [], thr

    def offer(self, score, idx):
        if not score >= self.f:                      # the kernel's `s >= L.f`
            return
        key = _key(score, idx)
        if len(self.keys) < self.k:
            self.keys.append(key)
        elif key > min(self.keys):
            self.keys[self.keys.index(min(self.keys))] = key
        else:
            return
        if len(self.keys) == self.k:
            self.f = min(self.keys)[0]

    def kth_score(self):
        return self.f if len(self.keys) == self.k else None


def symmetric_sweep_model(S, k, thr, B, rng, cap=None):
    """S: symmetric [n,n] scores with -inf where a pair is excluded (window / mask).  B: block = tile size.
    Returns (lists per keyframe as sorted keys, overflowed?)."""
    n = S.shape[0]
    nb = (n + B - 1) // B
    # every block's tiles [b, nb) cut into random runs; all runs of all blocks executed in a random interleaving
    runs = []
    for b in range(nb):
        cuts = sorted(set([b, nb] + [int(c) for c in rng.integers(b, nb + 1, size=rng.integers(0, 3))]))
        runs += [(b, lo, hi) for lo, hi in zip(cuts[:-1], cuts[1:]) if hi > lo]
    cursors = [[b, lo, hi, None] for b, lo, hi in runs]                 # [block, next tile, end, per-row lists]
    published = [[] for _ in range(n)]                                  # history of published bounds (monotone)
    buf = [[] for _ in range(n)]
    flushed = [[] for _ in range(n)]
    overflow = False
    live = list(range(len(cursors)))
    while live:
        ci = live[int(rng.integers(len(live)))]
        cur = cursors[ci]
        b, t = cur[0], cur[1]
        rows = range(b * B, min(n, (b + 1) * B))
        cols = range(t * B, min(n, (t + 1) * B))
        if cur[3] is None:
            cur[3] = {r: RunList(k, thr) for r in rows}
        # bounds staged for this tile: any value a column's owner has published so far, or none yet (threshold)
        staged = {}
        for c in cols:
            hist = published[c]
            staged[c] = thr if (not hist or rng.random() < 0.3) else hist[int(rng.integers(len(hist)))]
        for r in rows:
            for c in cols:
                s = S[r, c]
                cur[3][r].offer(s, c)                                   # row direction
                if t != b and s >= staged[c] and np.isfinite(s):        # column direction (never on the diagonal tile)
                    buf[c].append(_key(s, r))
                    if cap is not None and len(buf[c]) > cap:
                        overflow = True
        for r in rows:                                                  # publish (atomicMax: monotone)
            kth = cur[3][r].kth_score()
            if kth is not None and (not published[r] or kth > published[r][-1]):
                published[r].append(kth)
        cur[1] += 1
        if cur[1] == cur[2]:
            for r in rows:
                flushed[r] += cur[3][r].keys
            live.remove(ci)
    out = []
    for q in range(n):
        keys = sorted(set(flushed[q] + (buf[q][:cap] if cap is not None else buf[q])), reverse=True)[:k]
        out.append(keys)
    return out, overflow


def brute_force(S, k, thr):
    out = []
    for q in range(S.shape[0]):
        keys = [_key(S[q, j], j) for j in range(S.shape[1]) if S[q, j] >= thr and np.isfinite(S[q, j])]
        out.append(sorted(keys, reverse=True)[:k])
    return out


@pytest.mark.parametrize("seed", range(12))
def test_symmetric_bookkeeping_equals_brute_force(seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(20, 90))
    B = int(rng.choice([4, 8, 16]))
    k = int(rng.choice([1, 3, 5, 8]))
    x = rng.standard_normal((n, 6)).astype(np.float32)
    if seed % 3 == 0:
        x[rng.integers(0, n, size=n // 3)] = x[0]          # many exact ties: the index tie-break must hold
    xn = O.l2_normalize(x)
    S = (xn @ xn.T).astype(np.float32)
    S = np.maximum(S, S.T)                                 # exactly symmetric, as one computed tile is
    ts = np.arange(n) * 0.5
    S[O.time_excluded(ts, ts, float(rng.choice([0.0, 1.0, 2.6])))] = -np.inf
    thr = np.float32(rng.choice([-np.inf, -0.2, 0.3, 0.7]))
    got, _ = symmetric_sweep_model(S, k, thr, B, rng)
    want = brute_force(S, k, thr)
    assert got == want


def test_symmetric_bookkeeping_overflow_is_detected_not_silent():
    """With a buffer too small for a permissive threshold the model (like the kernel) must notice: the lists built
    from truncated buffers may be wrong, which is why the flag arms the full sweep."""
    rng = np.random.default_rng(5)
    n, B, k = 80, 8, 4
    x = rng.standard_normal((n, 5)).astype(np.float32)
    xn = O.l2_normalize(x)
    S = (xn @ xn.T).astype(np.float32)
    S = np.maximum(S, S.T)
    got, overflow = symmetric_sweep_model(S, k, np.float32(-np.inf), B, rng, cap=6)
    assert overflow
    got2, overflow2 = symmetric_sweep_model(S, k, np.float32(-np.inf), B, rng, cap=10 ** 6)
    assert not overflow2 and got2 == brute_force(S, k, np.float32(-np.inf))
