"""TEST INFRASTRUCTURE ONLY — CPU oracle for gated loop-closure retrieval.

A vectorised numpy restatement of the reference's algorithm for the hot path
(`scripts/semantic_gating/place_recognition.py` -> `loop_closure_gate.py` in the
reference repository; citations below are relative to that directory's parent
`/root/reference/scripts/semantic_gating/`).  The reference itself is an
O(N^2) Python loop over an N x N matrix and cannot run beyond ~10^4 keyframes;
this file computes the same decisions blocked over query rows.

Pinning status
  * floor gate (integer decisions): PINNED by the reference's published counts
    (`results/semantic_gating/lego_loam_semantic_analysis.txt:20-22`,
    `orb_slam3_semantic_analysis.txt:20-22`) — see tests/test_oracle_golden.py.
  * similarity / temporal mask / top-k: the reference ships no golden vectors
    or tests for them; PINNED instead against outputs of the unmodified
    reference code run in the build container (`tests/golden/make_golden.py`
    -> `tests/golden/*.npz`) and, when /root/reference is present, live in
    tests/test_oracle_golden.py::test_live_reference_random_cases.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.  The product package never does.
"""
from __future__ import annotations

import numpy as np

FLOOR_NONE = np.int32(-2**31)  # encodes `floor_label=None` (place_recognition.py:78,898)

GATE_FLAG = 0   # reference order: top-k first, floor check only flags (place_recognition.py:888-899)
GATE_MASK = 1   # fused-gate variant: cross-floor columns are excluded before top-k


# ----------------------------------------------------------------------------
# building blocks
# ----------------------------------------------------------------------------
def l2_normalize(x: np.ndarray) -> np.ndarray:
    """`x / (||x|| + 1e-8)` row-wise, in x's dtype (place_recognition.py:186-187, :169-170)."""
    x = np.asarray(x)
    if x.ndim == 1:
        return x / (np.linalg.norm(x) + 1e-8)
    norms = np.linalg.norm(x, axis=1, keepdims=True)
    return x / (norms + 1e-8)


def bf16_round(x: np.ndarray) -> np.ndarray:
    """Round fp32 to the nearest bf16 (ties to even), returned as fp32.
    Models the GPU path's operand precision (bf16 in, fp32 accumulate)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    rounding = np.uint64(0x7FFF) + ((u >> np.uint64(16)) & np.uint64(1))
    r = ((u + rounding) >> np.uint64(16)) << np.uint64(16)
    nan = np.isnan(x)
    out = r.astype(np.uint32).view(np.float32).reshape(x.shape)
    if nan.any():
        out = out.copy()
        out[nan] = np.nan
    return out


def encode_floors(floors) -> np.ndarray:
    """Python floor labels (ints or None) -> int32 with FLOOR_NONE for None."""
    if floors is None:
        return None
    out = np.empty(len(floors), dtype=np.int32)
    for i, f in enumerate(floors):
        out[i] = FLOOR_NONE if f is None else int(f)
    return out


def floor_ok(qf: np.ndarray, mf: np.ndarray, max_floor_diff: int) -> np.ndarray:
    """Gate predicate.  max_floor_diff: -1 gating off; 0 strict (reject any
    difference, loop_closure_gate.py:91 and place_recognition.py:898-899);
    1 non-strict (reject if |diff| > 1, loop_closure_gate.py:95).
    A FLOOR_NONE label on either side always passes (place_recognition.py:898)."""
    qf = np.asarray(qf, dtype=np.int64)
    mf = np.asarray(mf, dtype=np.int64)
    if max_floor_diff < 0:
        return np.ones(np.broadcast(qf, mf).shape, dtype=bool)
    none = (qf == int(FLOOR_NONE)) | (mf == int(FLOOR_NONE))
    return none | (np.abs(qf - mf) <= max_floor_diff)


def time_excluded(q_ts: np.ndarray, db_ts: np.ndarray, min_time_gap: float) -> np.ndarray:
    """`abs(t_db - t_q) < min_time_gap` in fp64, strict (place_recognition.py:884, :146).
    Returns bool [Q, N]."""
    q_ts = np.asarray(q_ts, dtype=np.float64)
    db_ts = np.asarray(db_ts, dtype=np.float64)
    return np.abs(db_ts[None, :] - q_ts[:, None]) < float(min_time_gap)


def _topk_rows(S: np.ndarray, k: int):
    """Per-row top-k of S under the total order (score desc, index asc).
    Entries equal to -inf are never returned.  Returns (scores[Q,k] -inf padded,
    idx[Q,k] -1 padded, count[Q])."""
    Q, N = S.shape
    kk = min(k, N)
    scores = np.full((Q, k), -np.inf, dtype=S.dtype)
    idx = np.full((Q, k), -1, dtype=np.int64)
    count = np.zeros(Q, dtype=np.int32)
    if kk == 0:
        return scores, idx, count
    if kk < N:
        part = np.argpartition(-S, kk - 1, axis=1)[:, :kk]
    else:
        part = np.broadcast_to(np.arange(N), (Q, N)).copy()
    pv = np.take_along_axis(S, part, axis=1)
    kth = pv.min(axis=1)
    # boundary ties: the partition picked arbitrary members of the tie class
    n_ge = (S >= kth[:, None]).sum(axis=1)
    for r in range(Q):
        if n_ge[r] > kk and np.isfinite(kth[r]):
            cand = np.nonzero(S[r] >= kth[r])[0]
            order = np.lexsort((cand, -S[r, cand]))[:kk]
            part[r] = cand[order]
            pv[r] = S[r, part[r]]
    # two stable passes = lexicographic (score desc, index asc)
    o = np.argsort(part, axis=1, kind="stable")
    part = np.take_along_axis(part, o, axis=1)
    pv = np.take_along_axis(pv, o, axis=1)
    o = np.argsort(-pv, axis=1, kind="stable")
    part = np.take_along_axis(part, o, axis=1)
    pv = np.take_along_axis(pv, o, axis=1)
    keep = pv > -np.inf                      # -inf sorts last, so kept entries are a prefix
    scores[:, :kk] = np.where(keep, pv, -np.inf)
    idx[:, :kk] = np.where(keep, part, -1)
    count[:] = keep.sum(axis=1)
    return scores, idx, count


# ----------------------------------------------------------------------------
# the fused path, restated
# ----------------------------------------------------------------------------
def gated_topk(q, db, q_ts=None, db_ts=None, q_floor=None, db_floor=None, *, k=10,
               threshold=-np.inf, min_time_gap=10.0, max_floor_diff=0,
               gate_mode=GATE_FLAG, normalize=True, bf16=False, block=1024,
               db_index_offset=0):
    """Top-k gated retrieval of every row of `q` against `db`.

    Follows find_loop_closures (place_recognition.py:868-909): similarities of
    row-normalised descriptors; temporal exclusion `abs(t_j - t_i) < gap` when
    timestamps are given; top-k; drop scores `< threshold` (compared in the
    score dtype: numpy's weak-scalar rule casts the Python float threshold to
    fp32 for fp32 descriptors); floor check flags (`GATE_FLAG`) or masks
    (`GATE_MASK`).  `bf16=True` rounds the normalised operands to bf16 first
    (the GPU path's arithmetic) — used to check kernel indexing tightly.

    Returns dict(scores[Q,k], idx[Q,k] int64 (global = local + db_index_offset),
    valid[Q,k] bool, count[Q] int32).
    """
    q = np.asarray(q)
    db = np.asarray(db)
    Q, N = q.shape[0], db.shape[0]
    qn = l2_normalize(q) if normalize else q
    dbn = l2_normalize(db) if normalize else db
    if bf16:
        qn, dbn = bf16_round(qn), bf16_round(dbn)
    sdtype = np.result_type(qn.dtype, dbn.dtype)
    thr = np.asarray(threshold).astype(sdtype)
    use_time = q_ts is not None and db_ts is not None
    qf = None if q_floor is None else np.asarray(q_floor, dtype=np.int64)
    mf = None if db_floor is None else np.asarray(db_floor, dtype=np.int64)
    gating = qf is not None and mf is not None and max_floor_diff >= 0

    scores = np.full((Q, k), -np.inf, dtype=sdtype)
    idx = np.full((Q, k), -1, dtype=np.int64)
    valid = np.zeros((Q, k), dtype=bool)
    count = np.zeros(Q, dtype=np.int32)
    dbT = np.ascontiguousarray(dbn.T)
    for s in range(0, Q, block):
        e = min(Q, s + block)
        S = qn[s:e] @ dbT
        if use_time:
            S[time_excluded(q_ts[s:e], db_ts, min_time_gap)] = -np.inf
        if gating and gate_mode == GATE_MASK:
            S[~floor_ok(qf[s:e, None], mf[None, :], max_floor_diff)] = -np.inf
        S[S < thr] = -np.inf          # threshold commutes with top-k (place_recognition.py:888-892)
        sc, ix, ct = _topk_rows(S, k)
        scores[s:e], idx[s:e], count[s:e] = sc, ix, ct
    got = idx >= 0
    if gating:
        mfl = mf[np.where(got, idx, 0)]
        valid = got & floor_ok(qf[:, None], mfl, max_floor_diff)
    else:
        valid = got.copy()
    idx = np.where(got, idx + db_index_offset, -1)
    return dict(scores=scores, idx=idx, valid=valid, count=count)


def compact(res):
    """Padded [Q,k] lists -> flat candidate arrays in the reference's order
    (query ascending, score descending; place_recognition.py:873,888)."""
    got = res["idx"] >= 0
    q_idx = np.nonzero(got)[0].astype(np.int64)
    return dict(query_idx=q_idx, match_idx=res["idx"][got], similarity=res["scores"][got],
                is_valid=res["valid"][got])


def find_loop_closures(desc, ts, floors, *, similarity_threshold=0.5, min_time_gap=10.0,
                       k=10, enable_floor_gating=True, gate_mode=GATE_FLAG, bf16=False):
    """SemanticPlaceRecognition.find_loop_closures (place_recognition.py:851-911)
    as flat arrays.  `floors`: int32 array with FLOOR_NONE for None, or None."""
    desc = np.asarray(desc)
    if desc.shape[0] < 2:                                   # place_recognition.py:864
        e = np.zeros(0, dtype=np.int64)
        return dict(query_idx=e, match_idx=e.copy(), similarity=np.zeros(0, np.float32),
                    is_valid=np.zeros(0, bool))
    res = gated_topk(desc, desc, ts, ts, floors, floors, k=k, threshold=similarity_threshold,
                     min_time_gap=min_time_gap,
                     max_floor_diff=0 if (enable_floor_gating and floors is not None) else -1,
                     gate_mode=gate_mode, bf16=bf16)
    return compact(res)


def query(q_desc, db, timestamp=None, db_ts=None, *, k=5, min_time_gap=10.0, bf16=False):
    """BasePlaceRecognition.query (place_recognition.py:117-163): no threshold,
    no floor check; temporal mask only when a timestamp is given (:144);
    masked entries are dropped (:154)."""
    q = np.asarray(q_desc)[None, :]
    q_ts = None if timestamp is None else np.asarray([timestamp], dtype=np.float64)
    res = gated_topk(q, db, q_ts, None if timestamp is None else db_ts, None, None, k=k,
                     threshold=-np.inf, min_time_gap=min_time_gap, max_floor_diff=-1, bf16=bf16)
    c = int(res["count"][0])
    return res["idx"][0, :c], res["scores"][0, :c]


def gate_candidates(floor_labels, query_idx, match_idx, strict_mode=True):
    """SemanticLoopClosureGate.gate_candidates (loop_closure_gate.py:60-126) on
    index arrays: returns (is_valid bool[M], stats dict with the reference's
    counter names, :53-58)."""
    fl = np.asarray(floor_labels, dtype=np.int64)
    qf = fl[np.asarray(query_idx, dtype=np.int64)]
    mf = fl[np.asarray(match_idx, dtype=np.int64)]
    diff = np.abs(qf - mf)
    ok = (diff == 0) if strict_mode else (diff <= 1)
    stats = {"total_candidates": int(ok.size), "accepted": int(ok.sum()),
             "rejected_cross_floor": int(ok.size - ok.sum()), "rejected_other": 0}
    return ok, stats


def spatial_candidates(positions, distance_threshold=2.0, min_index_gap=100):
    """detect_loop_closure_candidates (orb_slam3_integration.py:167-217): pairs
    i<j with ||p_i - p_j|| <= r and |i-j| >= gap.  Returns (i[M], j[M]) sorted
    by (i, j) like the reference's nested loops over sorted neighbour lists
    (order is irrelevant for the published counts)."""
    from scipy.spatial import cKDTree
    p = np.asarray(positions, dtype=np.float64)
    pairs = cKDTree(p).query_pairs(distance_threshold, output_type="ndarray")
    i, j = pairs[:, 0], pairs[:, 1]
    keep = (j - i) >= min_index_gap
    i, j = i[keep], j[keep]
    o = np.lexsort((j, i))
    return i[o].astype(np.int64), j[o].astype(np.int64)


def statistics(is_valid, similarity):
    """get_statistics (place_recognition.py:913-933)."""
    n = int(len(is_valid))
    if n == 0:
        return {"total_matches": 0, "valid_matches": 0, "rejected_matches": 0, "rejection_rate": 0.0}
    v = int(np.sum(is_valid))
    sim = np.asarray(similarity, dtype=np.float64)
    return {"total_matches": n, "valid_matches": v, "rejected_matches": n - v,
            "rejection_rate": (n - v) / n, "mean_similarity": float(sim.mean()),
            "mean_valid_similarity": float(sim[np.asarray(is_valid, bool)].mean()) if v > 0 else 0.0}


# ----------------------------------------------------------------------------
# candidate keys (the wire format of the multi-GPU all-gather; csrc/common.cuh)
# ----------------------------------------------------------------------------
def pack_keys(scores, idx):
    """(fp32 score, global index) -> int64 key whose unsigned order is (score desc, index asc);
    empty slots (idx < 0) -> 0.  Mirrors pack_key() in csrc/common.cuh."""
    s = np.ascontiguousarray(scores, dtype=np.float32)
    u = s.view(np.uint32).astype(np.uint64)
    neg = (u & np.uint64(0x80000000)) != 0
    o = np.where(neg, (~u) & np.uint64(0xFFFFFFFF), u | np.uint64(0x80000000))
    ix = np.asarray(idx, dtype=np.int64)
    key = (o << np.uint64(32)) | (np.uint64(0xFFFFFFFF) - np.where(ix >= 0, ix, 0).astype(np.uint64))
    key = np.where(ix >= 0, key, np.uint64(0))
    return key.view(np.int64)


def unpack_keys(keys):
    """int64 keys -> (scores fp32 with -inf for empty, idx int64 with -1 for empty)."""
    k = np.ascontiguousarray(keys).view(np.uint64)
    o = (k >> np.uint64(32)).astype(np.uint32)
    pos = (o & np.uint32(0x80000000)) != 0
    u = np.where(pos, o & np.uint32(0x7FFFFFFF), ~o)
    s = u.astype(np.uint32).view(np.float32)
    ix = (np.uint64(0xFFFFFFFF) - (k & np.uint64(0xFFFFFFFF))).astype(np.int64)
    empty = k == 0
    return np.where(empty, -np.inf, s).astype(np.float32), np.where(empty, -1, ix)


def merge_keys(keys_gathered, k):
    """[G,Q,k] keys -> [Q,k] largest keys per row, descending, 0 padded (K3's merge)."""
    g = np.ascontiguousarray(keys_gathered).view(np.uint64)
    G, Q, kk = g.shape
    allk = np.transpose(g, (1, 0, 2)).reshape(Q, G * kk)
    srt = np.sort(allk, axis=1)[:, ::-1][:, :k]
    return np.ascontiguousarray(srt).view(np.int64)


# ----------------------------------------------------------------------------
# CricaVPR cross-correlation re-rank (place_recognition.py:669-757)
# ----------------------------------------------------------------------------
def cross_correlation_score(query_features, match_features, bf16=False):
    """compute_cross_correlation_score (place_recognition.py:669-710): L2-normalise the patch
    rows (`x / (||x|| + 1e-8)`), correlation = q m^T, mean of the row maxima times mean of the
    column maxima, square root.  fp32 like the reference's torch-CPU path."""
    q = np.asarray(query_features, dtype=np.float32)
    m = np.asarray(match_features, dtype=np.float32)
    if q.ndim == 3:
        q = q[0]
    if m.ndim == 3:
        m = m[0]
    q = q / (np.linalg.norm(q, axis=-1, keepdims=True) + np.float32(1e-8))
    m = m / (np.linalg.norm(m, axis=-1, keepdims=True) + np.float32(1e-8))
    if bf16:
        q, m = bf16_round(q), bf16_round(m)
    corr = q @ m.T
    with np.errstate(invalid="ignore"):
        return np.float32(np.sqrt(corr.max(axis=1).mean(dtype=np.float32) * corr.max(axis=0).mean(dtype=np.float32)))


def rerank_candidates(features, query_idx, candidates, top_k=5, use_reranking=True, bf16=False):
    """rerank_candidates (place_recognition.py:712-757).  `features`: dict index -> [P,D] (or
    [1,P,D]) local features; `candidates`: list of (match_idx, global_similarity).  Combined score
    0.5*global + 0.5*cross when the match has cached features, else the global score; stable sort
    descending; first top_k."""
    if not use_reranking or query_idx not in features:
        return list(candidates[:top_k])
    out = []
    for match_idx, global_sim in candidates:
        if match_idx in features:
            cc = float(cross_correlation_score(features[query_idx], features[match_idx], bf16=bf16))
            out.append((match_idx, 0.5 * global_sim + 0.5 * cc))
        else:
            out.append((match_idx, global_sim))
    out.sort(key=lambda x: x[1], reverse=True)
    return out[:top_k]
