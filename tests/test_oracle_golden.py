"""CPU tier: pin the oracle restatement against the golden vectors produced by the
unmodified reference (tests/golden/make_golden.py), and against the reference run
live when /root/reference is present (build container only)."""
import glob
import os

import numpy as np
import pytest

from oracle import ref_loader
from oracle import semgate_oracle as O
import parity

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FLC = sorted(glob.glob(os.path.join(GOLDEN, "flc_*.npz")))
QRY = sorted(glob.glob(os.path.join(GOLDEN, "query_*.npz")))


def test_golden_present():
    assert len(FLC) >= 7 and len(QRY) >= 3
    assert os.path.isfile(os.path.join(GOLDEN, "gate_lego_loam.npz"))
    assert os.path.isfile(os.path.join(GOLDEN, "gate_orb_slam3.npz"))


@pytest.mark.parametrize("path", FLC, ids=[os.path.basename(p)[:-4] for p in FLC])
def test_find_loop_closures_matches_reference(path):
    c = parity.load_flc_case(path)
    got = O.find_loop_closures(c["desc"], c["ts"], O.encode_floors(c["floors"]),
                               similarity_threshold=c["thr"], min_time_gap=c["gap"], k=c["k"],
                               enable_floor_gating=c["gating"])
    ref = c["ref"]
    parity.check_order(got)
    parity.check_order(ref)
    # fp32 sgemm on both sides: same BLAS -> expect identical sets and ~1e-7 scores
    rep = parity.compare_candidates(ref, got, c["k"], c["thr"], tol=1e-5)
    assert rep["boundary_diffs"] == 0
    assert len(got["query_idx"]) == len(ref["query_idx"])
    assert np.array_equal(got["query_idx"], ref["query_idx"])
    enc = O.encode_floors(c["floors"])
    parity.check_decisions_exact(got, c["ts"], enc if c["gating"] else None, c["gap"], 0)
    # timestamps carried on PlaceMatch (place_recognition.py:905-906)
    assert np.array_equal(c["ts"][ref["query_idx"]], ref["query_timestamp"])
    assert np.array_equal(c["ts"][ref["match_idx"]], ref["match_timestamp"])


@pytest.mark.parametrize("path", QRY, ids=[os.path.basename(p)[:-4] for p in QRY])
def test_query_matches_reference(path):
    from semgate import synthetic
    g = np.load(path)
    n, d, seed, k, with_ts = [int(v) for v in g["params"]]
    gap = float(g["fparams"][0])
    desc, ts, _ = synthetic.make_case(n + 4, d, 3, seed, 0.5)
    for r, qi in enumerate(range(n, n + 4)):
        tq = float(ts[(qi * 97) % n]) + 0.25 if with_ts else None
        idx, sim = O.query(desc[qi], desc[:n], tq, ts[:n], k=k, min_time_gap=gap)
        cnt = int(g["count"][r])
        assert len(idx) == cnt
        assert np.array_equal(idx, g["match_idx"][r, :cnt])
        assert np.allclose(sim, g["similarity"][r, :cnt], atol=1e-5)


@pytest.mark.parametrize("algo", ["lego_loam", "orb_slam3"])
def test_gate_published_counts(algo):
    """Integer KAT: results/semantic_gating/{algo}_semantic_analysis.txt:20-22."""
    g = np.load(os.path.join(GOLDEN, f"gate_{algo}.npz"))
    i, j = O.spatial_candidates(g["positions"], 2.0, 100)
    ok, stats = O.gate_candidates(g["floor_labels"], i, j, strict_mode=True)
    total, acc, rej = [int(v) for v in g["published"]]
    assert (stats["total_candidates"], stats["accepted"], stats["rejected_cross_floor"]) == (total, acc, rej)
    ok2, stats2 = O.gate_candidates(g["floor_labels"], i, j, strict_mode=False)
    assert (stats2["accepted"], stats2["rejected_cross_floor"]) == tuple(int(v) for v in g["nonstrict"])


def test_bf16_round_matches_torch():
    import torch
    x = np.random.default_rng(0).standard_normal(100000).astype(np.float32) * 3
    x[:4] = [0.0, -0.0, 1.0, 65504.0]
    want = torch.from_numpy(x).to(torch.bfloat16).to(torch.float32).numpy()
    assert np.array_equal(O.bf16_round(x), want)


def test_bf16_model_within_tolerance():
    """The GPU arithmetic model (bf16 operands, fp32 accumulate) stays within the
    north-star 2e-3 of the fp32 reference path on the synthetic distribution."""
    from semgate import synthetic
    desc, ts, fl = synthetic.make_case(600, 512, 3, 21)
    a = O.gated_topk(desc, desc, ts, ts, fl, fl, k=25, threshold=0.5)
    b = O.gated_topk(desc, desc, ts, ts, fl, fl, k=25, threshold=0.5, bf16=True)
    rep = parity.compare_candidates(O.compact(a), O.compact(b), 25, 0.5)
    assert rep["max_score_err"] < parity.SCORE_TOL


def test_edge_cases():
    from semgate import synthetic
    desc, ts, fl = synthetic.make_case(50, 32, 3, 5)
    # N < 2 -> [] (place_recognition.py:864)
    assert len(O.find_loop_closures(desc[:1], ts[:1], fl[:1])["query_idx"]) == 0
    assert len(O.find_loop_closures(desc[:0], ts[:0], fl[:0])["query_idx"]) == 0
    # k larger than the database
    r = O.find_loop_closures(desc[:6], ts[:6], fl[:6], similarity_threshold=-1.0, min_time_gap=0.6, k=25)
    assert np.bincount(r["query_idx"], minlength=6).max() <= 5
    # gap = 0 keeps the self pair (abs(0) < 0 is False, place_recognition.py:884)
    r = O.find_loop_closures(desc, ts, fl, similarity_threshold=0.99, min_time_gap=0.0, k=3)
    assert np.array_equal(r["query_idx"], r["match_idx"])
    # mask mode is a superset of flag mode on the valid set
    a = O.compact(O.gated_topk(desc, desc, ts, ts, fl, fl, k=3, threshold=0.2, gate_mode=O.GATE_FLAG))
    b = O.compact(O.gated_topk(desc, desc, ts, ts, fl, fl, k=3, threshold=0.2, gate_mode=O.GATE_MASK))
    va = set(zip(a["query_idx"][a["is_valid"]].tolist(), a["match_idx"][a["is_valid"]].tolist()))
    vb = set(zip(b["query_idx"].tolist(), b["match_idx"].tolist()))
    assert b["is_valid"].all() and va <= vb


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present on this machine")
def test_live_reference_random_cases():
    """Build container only: fresh seeds through the unmodified reference."""
    from semgate import synthetic
    for seed, n, d, k, thr, gap in [(101, 180, 48, 4, 0.45, 3.0), (102, 220, 80, 12, 0.3, 7.5)]:
        desc, ts, fl = synthetic.make_case(n, d, 4, seed, 0.5)
        _, ms = ref_loader.run_find_loop_closures(desc, ts, [int(f) for f in fl], thr, gap, k, True)
        ref = dict(query_idx=np.array([m.query_idx for m in ms]), match_idx=np.array([m.match_idx for m in ms]),
                   similarity=np.array([m.similarity for m in ms]), is_valid=np.array([m.is_valid for m in ms]))
        got = O.find_loop_closures(desc, ts, fl.astype(np.int32), similarity_threshold=thr, min_time_gap=gap, k=k)
        rep = parity.compare_candidates(ref, got, k, thr, tol=1e-5)
        assert rep["boundary_diffs"] == 0


RER = sorted(glob.glob(os.path.join(GOLDEN, "rerank_*.npz")))


@pytest.mark.parametrize("path", RER, ids=[os.path.basename(p)[:-4] for p in RER])
def test_rerank_matches_reference(path):
    """CricaVPR cross-correlation re-rank (place_recognition.py:669-757) against the reference's own outputs."""
    from semgate import synthetic
    g = np.load(path)
    n, patches, dim, seed, top_k, ncand, missing = [int(v) for v in g["params"]]
    feats, _ = synthetic.make_local_features(n, patches, dim, seed)
    cache = {i: feats[i] for i in range(n) if i != missing}
    for r, q in enumerate(g["query_idx"].tolist()):
        cands = [(int(c), float(s)) for c, s in zip(g["cand_idx"][r], g["cand_sim"][r])]
        for ci, (c, _) in enumerate(cands):
            if c != missing:
                assert abs(float(O.cross_correlation_score(feats[q], feats[c])) - float(g["cross"][r, ci])) < 2e-6
        rr = O.rerank_candidates(cache, q, cands, top_k=top_k)
        cnt = int(g["out_count"][r])
        assert len(rr) == cnt
        assert [m for m, _ in rr] == g["out_idx"][r, :cnt].tolist()
        assert np.allclose([s for _, s in rr], g["out_score"][r, :cnt], atol=2e-6)
    # the bf16 model of the GPU arithmetic stays inside the 2e-3 tolerance
    a = O.cross_correlation_score(feats[0], feats[1])
    b = O.cross_correlation_score(feats[0], feats[1], bf16=True)
    assert abs(float(a) - float(b)) < 2e-3
    # no cached query features / re-ranking off -> the input order is kept (place_recognition.py:733-737)
    assert O.rerank_candidates({}, 0, [(1, 0.5), (2, 0.9)], top_k=1) == [(1, 0.5)]
    assert O.rerank_candidates(cache, 0, [(1, 0.5), (2, 0.9)], top_k=5, use_reranking=False) == [(1, 0.5), (2, 0.9)]
