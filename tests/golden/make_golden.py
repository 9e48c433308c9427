"""Generate golden vectors by running the UNMODIFIED reference in the build
container (needs /root/reference; the GPU box does not have it, which is why
the outputs are committed).

    python tests/golden/make_golden.py

Writes (all small, compressed):
  flc_*.npz     find_loop_closures outputs (place_recognition.py:851-911) on seeded
                synthetic cases: inputs are regenerated from (n, d, floors, seed) by
                semgate.synthetic, outputs are the reference PlaceMatch fields.
  query_*.npz   BasePlaceRecognition.query outputs (place_recognition.py:117-163).
  rerank_*.npz  CricaVPR cross-correlation scores and re-ranked lists (place_recognition.py:669-757).
  gate_lego_loam.npz / gate_orb_slam3.npz
                positions + file-membership floor labels of the shipped trajectories
                (results/trajectories/{lego_loam,orb_slam3}/*.txt; label order 5,1,4,2 from
                lego_loam_integration.py:55-60 / orb_slam3_integration.py:58-63) and the
                counts the reference publishes
                (results/semantic_gating/*_semantic_analysis.txt:20-22), re-derived here
                through the reference's own SemanticLoopClosureGate.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "multi-level-indoor-slam_b200"))

from oracle import ref_loader  # noqa: E402
from semgate import synthetic  # noqa: E402

# (name, n, d, num_floors, seed, dt, threshold, gap, k, gating, none_every)
FLC_CASES = [
    ("small",      300,  64, 3, 1, 0.5, 0.5, 10.0, 10, True, 0),
    ("k25",        700, 128, 3, 2, 0.5, 0.5, 10.0, 25, True, 0),
    ("nogate",     400,  96, 4, 3, 0.5, 0.5, 10.0,  5, False, 0),
    ("none_floor", 350,  64, 3, 4, 0.5, 0.4,  4.0,  8, True, 7),
    ("gap0",       256,  64, 3, 5, 0.5, 0.5,  0.0,  6, True, 0),
    ("lowthr",     200,  64, 3, 6, 2.0, -1.0, 10.0, 25, True, 0),
    ("c1_slice",  1000, 512, 3, 0, 0.5, 0.5, 10.0, 25, True, 0),
]

# (name, n, d, seed, k, gap, with_timestamp)
QUERY_CASES = [
    ("ts",   500, 128, 11, 5, 10.0, True),
    ("nots", 500, 128, 12, 7, 10.0, False),
    ("k25",  900, 256, 13, 25, 3.0, True),
]


def gen_flc():
    for name, n, d, nf, seed, dt, thr, gap, k, gating, none_every in FLC_CASES:
        desc, ts, floors = synthetic.make_case(n, d, nf, seed, dt)
        fl = [None if (none_every and i % none_every == 0) else int(f) for i, f in enumerate(floors)]
        _, matches = ref_loader.run_find_loop_closures(desc, ts, fl, thr, gap, k, gating)
        np.savez_compressed(
            os.path.join(HERE, f"flc_{name}.npz"),
            params=np.array([n, d, nf, seed, k, none_every, int(gating)], dtype=np.int64),
            fparams=np.array([dt, thr, gap], dtype=np.float64),
            query_idx=np.array([m.query_idx for m in matches], dtype=np.int32),
            match_idx=np.array([m.match_idx for m in matches], dtype=np.int32),
            similarity=np.array([m.similarity for m in matches], dtype=np.float32),
            is_valid=np.array([m.is_valid for m in matches], dtype=bool),
            query_timestamp=np.array([m.query_timestamp for m in matches], dtype=np.float64),
            match_timestamp=np.array([m.match_timestamp for m in matches], dtype=np.float64))
        print(f"flc_{name}: {len(matches)} matches, {sum(m.is_valid for m in matches)} valid")


def gen_query():
    for name, n, d, seed, k, gap, with_ts in QUERY_CASES:
        desc, ts, floors = synthetic.make_case(n + 4, d, 3, seed, 0.5)
        vpr = ref_loader.identity_vpr(d)
        for i in range(n):
            vpr.add_image(desc[i], float(ts[i]), int(floors[i]))
        out_idx, out_sim, out_cnt = [], [], []
        for qi in range(n, n + 4):
            # query timestamps placed inside the DB's time range so the window bites
            tq = float(ts[(qi * 97) % n]) + 0.25 if with_ts else None
            ms = vpr.query(desc[qi], tq, k=k, min_time_gap=gap)
            assert all(m.query_idx == n for m in ms)
            out_cnt.append(len(ms))
            out_idx.append([m.match_idx for m in ms] + [-1] * (k - len(ms)))
            out_sim.append([m.similarity for m in ms] + [-np.inf] * (k - len(ms)))
        np.savez_compressed(
            os.path.join(HERE, f"query_{name}.npz"),
            params=np.array([n, d, seed, k, int(with_ts)], dtype=np.int64),
            fparams=np.array([gap], dtype=np.float64),
            match_idx=np.array(out_idx, dtype=np.int32),
            similarity=np.array(out_sim, dtype=np.float32),
            count=np.array(out_cnt, dtype=np.int32))
        print(f"query_{name}: counts {out_cnt}")


SEQ = [("5th_floor", 5), ("1st_floor", 1), ("4th_floor", 4), ("2nd_floor", 2)]
PUBLISHED = {  # results/semantic_gating/{algo}_semantic_analysis.txt:20-22
    "lego_loam": (87044, 21477, 65567),
    "orb_slam3": (5110618, 1498091, 3612527),
}


def gen_gate():
    LCG = ref_loader.loop_closure_gate()
    from scipy.spatial import cKDTree
    for algo, published in PUBLISHED.items():
        pos, lab = [], []
        for seq, floor in SEQ:
            path = os.path.join(ref_loader.REFERENCE_ROOT, "results", "trajectories", algo, f"{seq}.txt")
            traj = np.loadtxt(path)
            pos.append(traj[:, 1:4])
            lab.append(np.full(len(traj), floor, dtype=np.int64))
        pos = np.vstack(pos)
        lab = np.concatenate(lab)
        # candidate generation exactly as orb_slam3_integration.py:195-211 (ball query, |i-j|>=100, i<j)
        tree = cKDTree(pos)
        pairs = tree.query_pairs(2.0, output_type="ndarray")
        keep = (pairs[:, 1] - pairs[:, 0]) >= 100
        pairs = pairs[keep]
        total = len(pairs)
        if algo == "lego_loam":
            # the reference gate itself, candidate by candidate (loop_closure_gate.py:105-126)
            gate = LCG.SemanticLoopClosureGate(lab, strict_mode=True)
            valid, rejected = gate.gate_candidates([(int(i), int(j), 0.0) for i, j in pairs])
            acc, rej = len(valid), len(rejected)
            gate2 = LCG.SemanticLoopClosureGate(lab, strict_mode=False)
            v2, r2 = gate2.gate_candidates([(int(i), int(j), 0.0) for i, j in pairs])
            nonstrict = (len(v2), len(r2))
        else:
            same = lab[pairs[:, 0]] == lab[pairs[:, 1]]
            acc, rej = int(same.sum()), int((~same).sum())
            d = np.abs(lab[pairs[:, 0]] - lab[pairs[:, 1]])
            nonstrict = (int((d <= 1).sum()), int((d > 1).sum()))
        assert (total, acc, rej) == published, (algo, total, acc, rej, published)
        np.savez_compressed(
            os.path.join(HERE, f"gate_{algo}.npz"),
            positions=pos.astype(np.float64), floor_labels=lab.astype(np.int32),
            published=np.array(published, dtype=np.int64),
            nonstrict=np.array(nonstrict, dtype=np.int64))
        print(f"gate_{algo}: {total} / {acc} / {rej} == published; non-strict {nonstrict}")


# (name, n keyframes, patches, dim, seed, top_k, candidates per query, missing-feature index or -1)
RERANK_CASES = [
    ("small", 24, 48, 64, 3, 5, 10, -1),
    ("missing", 20, 33, 40, 4, 3, 8, 5),
    ("dinov2_shape", 10, 529, 768, 5, 5, 6, -1),
]


def gen_rerank():
    """CricaVPR.compute_cross_correlation_score / rerank_candidates (place_recognition.py:669-757) run
    verbatim (torch CPU).  The class constructor would load models; the two methods only touch
    `use_reranking` and `_feature_cache`."""
    PR = ref_loader.place_recognition()
    for name, n, patches, dim, seed, top_k, ncand, missing in RERANK_CASES:
        feats, _ = synthetic.make_local_features(n, patches, dim, seed)
        crica = PR.CricaVPR.__new__(PR.CricaVPR)
        crica.use_reranking = True
        crica._feature_cache = {i: feats[i][None] for i in range(n) if i != missing}   # [1,P,D] like extract_local_features
        rng = np.random.default_rng(seed + 100)
        q_idx, cand_idx, cand_sim, cross, out_idx, out_score, out_cnt = [], [], [], [], [], [], []
        for q in range(0, n, 2):
            if q == missing:
                continue
            cands = [int(c) for c in rng.choice([i for i in range(n) if i != q], size=ncand, replace=False)]
            sims = rng.uniform(0.3, 0.95, size=ncand).astype(np.float32)
            cl = [(c, float(s_)) for c, s_ in zip(cands, sims)]
            rr = crica.rerank_candidates(q, cl, top_k=top_k)
            q_idx.append(q)
            cand_idx.append(cands)
            cand_sim.append(sims)
            cross.append([crica.compute_cross_correlation_score(feats[q][None], feats[c][None]) if c != missing else np.nan
                          for c in cands])
            out_idx.append([m for m, _ in rr] + [-1] * (top_k - len(rr)))
            out_score.append([sc for _, sc in rr] + [np.nan] * (top_k - len(rr)))
            out_cnt.append(len(rr))
        np.savez_compressed(
            os.path.join(HERE, f"rerank_{name}.npz"),
            params=np.array([n, patches, dim, seed, top_k, ncand, missing], dtype=np.int64),
            query_idx=np.array(q_idx, dtype=np.int32), cand_idx=np.array(cand_idx, dtype=np.int32),
            cand_sim=np.array(cand_sim, dtype=np.float32), cross=np.array(cross, dtype=np.float32),
            out_idx=np.array(out_idx, dtype=np.int32), out_score=np.array(out_score, dtype=np.float64),
            out_count=np.array(out_cnt, dtype=np.int32))
        print(f"rerank_{name}: {len(q_idx)} queries x {ncand} candidates, cross in "
              f"[{np.nanmin(cross):.3f}, {np.nanmax(cross):.3f}]")


if __name__ == "__main__":
    if not ref_loader.available():
        sys.exit("reference not present; golden vectors can only be generated in the build container")
    only = sys.argv[1:] or ["flc", "query", "gate", "rerank"]
    if "flc" in only:
        gen_flc()
    if "query" in only:
        gen_query()
    if "gate" in only:
        gen_gate()
    if "rerank" in only:
        gen_rerank()
