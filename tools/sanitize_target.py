"""Small shapes of every kernel variant, for compute-sanitizer (one --tool per gpurun call):

    python tools/sanitize_target.py && compute-sanitizer --tool memcheck python tools/sanitize_target.py

Variants: K1 (block / cluster form), K2 <1,1>, <2,1>, <2,2>, symmetric <2,1,SYM> (run table, super-rows, overflow ->
armed full sweep, one part of a split), multi-pass k > 64, dense output, K3 (thread-per-row, block-per-row, peers'
row slice, sorted lists, the dense-list kernel), K4 (+ valid-only, row slice), statistics, gate, K5 re-rank (single-CTA and pair forms) + select,
K6 streaming query, spatial join, K2 with one / two epilogue sets, the one-call sweep (K3 -> K4 hand-off).
Every result is compared with the plain path where one exists, so a silent corruption also fails the run.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multi-level-indoor-slam_b200"))

import numpy as np
import torch

from semgate import _native, synthetic


def main():
    eng = _native.get_engine(0)
    dev = torch.device("cuda", 0)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    # K1
    for n, d in ((37, 100), (64, 4096), (5, 8448), (3, 49152)):
        eng.normalize_cast(torch.randn((n, d), device=dev))
    desc, ts, fl = synthetic.make_case(900, 128, 3, seed=1)
    fl = fl.astype(np.int32)
    xb = eng.normalize_cast(t(desc))
    tts, tfl = t(ts), t(fl)
    kw = dict(k=25, similarity_threshold=0.3, min_time_gap=5.0, max_floor_diff=0)
    full = eng.gated_topk(xb, xb, _native.make_params(symmetric=-1, cta_group=1, **kw), q_ts=tts, db_ts=tts, q_floor=tfl, db_floor=tfl, want_keys=True)
    for cg in (2, 4):
        r = eng.gated_topk(xb, xb, _native.make_params(symmetric=-1, cta_group=cg, **kw), q_ts=tts, db_ts=tts, q_floor=tfl, db_floor=tfl, want_keys=True)
        assert torch.equal(r.keys, full.keys), cg
    # rectangular, ragged, mask mode
    r = eng.gated_topk(xb[:300], xb[:700], _native.make_params(gate_mode=_native.GATE_MASK, cta_group=2, **kw), q_ts=tts[:300].contiguous(),
                       db_ts=tts[:700].contiguous(), q_floor=tfl[:300].contiguous(), db_floor=tfl[:700].contiguous())
    # symmetric: run table, super-row formula, one part of three, overflow -> armed full sweep
    for table in ("1", "0"):
        os.environ["SEMGATE_SYM_TABLE"] = table
        r = eng.gated_topk(xb, xb, _native.make_params(symmetric=1, cta_group=2, **kw), q_ts=tts, db_ts=tts, q_floor=tfl, db_floor=tfl, want_keys=True)
        assert eng.last_sweep_mode()[0] == 1 and torch.equal(r.keys, full.keys), table
        parts = [eng.gated_topk(xb, xb, _native.make_params(symmetric=1, cta_group=2, part_index=g, part_count=3, **kw), q_ts=tts, db_ts=tts,
                                q_floor=tfl, db_floor=tfl, want_keys=True, want_lists=False).keys for g in range(3)]
        m = eng.merge_topk(torch.stack(parts), 25, q_floor=tfl, db_floor_all=tfl, max_floor_diff=0, want_keys=True)
        assert torch.equal(m.keys, full.keys), ("parts", table)
    os.environ.pop("SEMGATE_SYM_TABLE")
    kw_all = dict(kw, similarity_threshold=-np.inf, min_time_gap=0.0)
    a = eng.gated_topk(xb, xb, _native.make_params(symmetric=1, cta_group=2, **kw_all), q_ts=tts, db_ts=tts, q_floor=tfl, db_floor=tfl, want_keys=True)
    mode = eng.last_sweep_mode()[0]
    b = eng.gated_topk(xb, xb, _native.make_params(symmetric=-1, cta_group=2, **kw_all), q_ts=tts, db_ts=tts, q_floor=tfl, db_floor=tfl, want_keys=True)
    assert torch.equal(a.keys, b.keys), "overflow path"
    print("symmetric sweep with an all-admitting threshold ran in mode", mode)
    # multi-pass k > 64, dense output
    r = eng.gated_topk(xb[:200], xb, _native.make_params(k=100, similarity_threshold=-0.05, min_time_gap=5.0, cta_group=1), q_ts=tts[:200].contiguous(), db_ts=tts)
    assert int(r.count.max()) > 64
    eng.similarity_matrix(xb[:130], xb[:300])
    # K6 + block-per-row K3
    big, bts, bfl = synthetic.make_case(6000, 256, 3, seed=2)
    bb = eng.normalize_cast(t(big))
    r6 = eng.gated_topk(bb[:2].contiguous(), bb, _native.make_params(k=25), q_ts=t(bts[:2]), db_ts=t(bts))
    r2 = eng.gated_topk(bb[:2].contiguous(), bb, _native.make_params(k=25, cta_group=1), q_ts=t(bts[:2]), db_ts=t(bts))
    assert torch.equal(r6.idx, r2.idx)
    # K3 peers' row slice + flags, K4 forms, statistics
    keys = torch.stack(parts)
    bufs = []
    for g in range(3):
        bbuf = torch.zeros((900 * 25 + 8,), dtype=torch.int64, device=dev)
        bbuf[:900 * 25] = keys[g].reshape(-1)
        bufs.append(bbuf)
    table_t = torch.tensor([p.data_ptr() for p in bufs], dtype=torch.int64, device=dev)
    flag = torch.zeros((1,), dtype=torch.int32, device=dev)
    mine = eng.merge_topk_peers_rows(table_t.data_ptr(), 3, 900, 25, 300, 300, q_floor=tfl, db_floor_all=tfl, max_floor_diff=0,
                                     want_keys=True, flag_offset=900 * 25, any_flag=flag)
    assert torch.equal(mine.keys, full.keys[300:600]) and int(flag.item()) == 0
    oq, om, os_, ov, tot = eng.compact(full)
    eng.compact(full, valid_only=True)
    eng.compact(mine, query_offset=300)
    eng.candidate_stats(os_, ov, tot)
    n_c = int(tot.item())
    eng.gate_candidates(tfl, oq[:n_c].contiguous(), om[:n_c].contiguous(), 0)
    # one-call sweep (eager form; the graph form is a replay of the same launches)
    eng.find_loop_closures_device(xb, _native.make_params(**kw), ts=tts, floor=tfl, use_graph=False)
    # K5 re-rank + select
    feats, _ = synthetic.make_local_features(6, 40, 64, seed=3)
    fb = eng.normalize_cast(t(feats.reshape(-1, 64))).view(6, 40, -1)
    qi = torch.tensor([0, 0, 1, 2, 5, -1], dtype=torch.int32, device=dev)
    mi = torch.tensor([1, 2, 3, 4, 0, 2], dtype=torch.int32, device=dev)
    cross, comb = eng.rerank_scores(fb, qi, mi, torch.rand(6, device=dev))
    eng.rerank_select(torch.arange(12, dtype=torch.int32, device=dev).view(2, 6), torch.rand((2, 6), device=dev),
                      torch.tensor([6, 4], dtype=torch.int32, device=dev), 3)
    # K5 pair form: ring + strip (P = 280), plain stages (ring forbidden), no strip (P = 300), against the single-CTA form
    for P in (280, 300):
        f2, _ = synthetic.make_local_features(8, P, 128, seed=P)
        fb2 = eng.normalize_cast(t(f2.reshape(-1, 128))).view(8, P, -1)
        q2 = torch.tensor([0, 0, 1, 2, 5, -1, 7, 3, 3], dtype=torch.int32, device=dev)
        m2 = torch.tensor([1, 2, 3, 4, 0, 2, 99, 6, 5], dtype=torch.int32, device=dev)
        g2 = torch.rand(9, device=dev)
        outs = []
        for env in ({}, {"SEMGATE_RERANK_RING": "0"}, {"SEMGATE_RERANK_UNROLL": "0"}, {"SEMGATE_RERANK_CLUSTER": "1"}):
            os.environ.update(env)
            outs.append(eng.rerank_scores(fb2, q2, m2, g2)[0].clone())
            for kk in env:
                os.environ.pop(kk)
        for o in outs[1:]:
            assert torch.allclose(o, outs[0], atol=2e-6, equal_nan=True), P
    # K2 with two epilogue sets (opt-in) against one, K3 network kernel on sorted lists
    os.environ["SEMGATE_EPI_SETS"] = "2"
    one = eng.gated_topk(xb, xb, _native.make_params(symmetric=-1, cta_group=2, **kw), q_ts=tts, db_ts=tts, q_floor=tfl, db_floor=tfl, want_keys=True)
    os.environ.pop("SEMGATE_EPI_SETS")
    assert torch.equal(one.keys, full.keys), "epilogue sets"
    ks = torch.sort(torch.randint(1, 2 ** 62, (3, 66000, 25), device=dev, dtype=torch.int64), dim=2, descending=True).values.contiguous()
    ms = eng.merge_topk(ks, 25, want_keys=True)
    want = torch.sort(ks.permute(1, 0, 2).reshape(66000, 75), dim=1, descending=True).values[:, :25]
    assert torch.equal(ms.keys, want), "sorted-list merge"
    # K3: that was the dense-list kernel (bulk copies); the network kernel on the same lists, a row slice that starts at an odd
    # row (plain loads), even k (pitched rows), unsorted lists (sorting network)
    eng.set_option("k3_dense", 0)
    assert torch.equal(eng.merge_topk(ks, 25, want_keys=True).keys, want), "network kernel"
    eng.set_option("k3_dense", 1)
    tab = torch.tensor([ks[g].data_ptr() for g in range(3)], dtype=torch.int64, device=dev)
    assert torch.equal(eng.merge_topk_peers_rows(tab.data_ptr(), 3, 66000, 25, 33, 65900, want_keys=True).keys, want[33:65933]), "dense rows"
    k10 = torch.randint(1, 2 ** 62, (2, 65600, 10), device=dev, dtype=torch.int64)
    w10 = torch.sort(k10.permute(1, 0, 2).reshape(65600, 20), dim=1, descending=True).values[:, :10]
    assert torch.equal(eng.merge_topk(k10, 10, want_keys=True).keys, w10), "dense, even k, unsorted"
    # pair gate with the label table in shared memory (>= 10 candidates per label)
    gl = torch.randint(1, 5, (3000,), dtype=torch.int32, device=dev)
    gq = torch.randint(-1, 3001, (40001,), dtype=torch.int32, device=dev)
    gm = torch.randint(0, 3000, (40001,), dtype=torch.int32, device=dev)
    gv, gc = eng.gate_candidates(gl, gq, gm, 0)
    okq = (gq >= 0) & (gq < 3000)
    wv = okq & (gl[gq.clamp(0, 2999).long()] == gl[gm.long()])
    assert torch.equal(gv.bool(), wv) and int(gc[2].item()) == int((~okq).sum().item()), "gate, shared-memory table"
    # spatial join
    pos = np.cumsum(np.random.default_rng(0).normal(size=(600, 3)) * 0.3, axis=0)
    eng.spatial_candidates_host(pos, 2.0, 50)
    torch.cuda.synchronize()
    print("sanitize_target ok")


if __name__ == "__main__":
    main()
