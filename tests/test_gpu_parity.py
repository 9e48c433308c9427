"""GPU tier (-m gpu): the sm_100a path, called through the C ABI, against the CPU oracle,
the golden vectors produced by the unmodified reference, and size-independent
properties at the benchmark's full size."""
import glob
import os

import numpy as np
import pytest

import parity
from oracle import semgate_oracle as O

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FLC = sorted(glob.glob(os.path.join(GOLDEN, "flc_*.npz")))
QRY = sorted(glob.glob(os.path.join(GOLDEN, "query_*.npz")))

# Against the oracle's model of the GPU arithmetic (bf16 operands, wide accumulation) the
# only differences are fp32 summation order and the rare bf16 rounding flip caused by the
# row norm being summed in a different order (one flipped element moves a score by about
# x_i * y_i * 2^-8 <= ~6e-5 at d=64, and a few can add up).  ~7x tighter than the north-star tolerance.
BF16_MODEL_TOL = 3e-4


@pytest.fixture(scope="module")
def eng():
    import torch
    assert torch.cuda.is_available(), "GPU tier needs a CUDA device"
    from semgate import _native
    return _native.get_engine(0)


def _t(a, dtype=None):
    import torch
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


def run_gpu(eng, q, db, k, thr=-np.inf, gap=10.0, q_ts=None, db_ts=None, q_fl=None, db_fl=None, mfd=-1,
            mode=0, cg=1, offset=0, sym=0):
    """normalise + gated_topk on the GPU; returns the padded result as numpy.  Arguments that are the same
    object on the host are the same tensor on the device (what the symmetric sweep keys on)."""
    import torch
    from semgate import _native
    qb = eng.normalize_cast(_t(q, torch.float32))
    dbb = qb if db is q else eng.normalize_cast(_t(db, torch.float32))
    p = _native.make_params(k=k, similarity_threshold=thr, min_time_gap=gap, max_floor_diff=mfd, gate_mode=mode,
                            db_index_offset=offset, cta_group=cg, symmetric=sym)
    tq = None if q_ts is None else _t(q_ts, torch.float64)
    td = tq if db_ts is q_ts else (None if db_ts is None else _t(db_ts, torch.float64))
    fq = None if q_fl is None else _t(q_fl, torch.int32)
    fd = fq if db_fl is q_fl else (None if db_fl is None else _t(db_fl, torch.int32))
    r = eng.gated_topk(qb, dbb, p, q_ts=tq, db_ts=td, q_floor=fq, db_floor=fd, want_keys=True)
    torch.cuda.synchronize()
    return dict(scores=r.scores.cpu().numpy(), idx=r.idx.cpu().numpy().astype(np.int64),
                valid=r.valid.cpu().numpy().astype(bool), count=r.count.cpu().numpy(), keys=r.keys.cpu().numpy(),
                mode=eng.last_sweep_mode()[0])


def check_padded(res, k):
    """Structural invariants of the padded lists."""
    sc, ix, ct = res["scores"], res["idx"], res["count"]
    Q = sc.shape[0]
    pos = np.arange(k)[None, :]
    filled = pos < ct[:, None]
    assert np.all(ix[filled] >= 0) and np.all(ix[~filled] == -1)
    assert np.all(np.isneginf(sc[~filled]))
    assert not np.any(res["valid"][~filled])
    d = np.diff(sc, axis=1)
    assert np.all(d[filled[:, 1:]] <= 0), "scores not descending"
    for r in range(min(Q, 512)):
        row = ix[r, :ct[r]]
        assert len(set(row.tolist())) == len(row), "duplicate database index in a list"


# --------------------------------------------------------------------------- K1
@pytest.mark.parametrize("n,d", [(1, 64), (37, 100), (300, 512), (64, 4096), (5, 8448), (3, 49152), (130, 33)])
def test_normalize_cast(eng, n, d):
    import torch
    rng = np.random.default_rng(n * 1000 + d)
    x = (rng.standard_normal((n, d)) * rng.uniform(0.1, 30, size=(n, 1))).astype(np.float32)
    if n > 2:
        x[1] = 0.0                                   # zero row: 0 / (0 + 1e-8) = 0
    out = eng.normalize_cast(_t(x))
    torch.cuda.synchronize()
    dp = ((d + 63) // 64) * 64
    assert out.shape == (n, dp) and out.dtype == torch.bfloat16
    got = out.float().cpu().numpy()
    want = O.bf16_round(O.l2_normalize(x))
    assert np.all(got[:, d:] == 0), "padding must be zero"
    # the row norm is summed in a different order than numpy's: allow one bf16 ulp
    ulp = np.maximum(np.abs(want), 2.0 ** -126) * 2.0 ** -7
    assert np.all(np.abs(got[:, :d] - want) <= ulp + 1e-30)
    assert np.mean(got[:, :d] == want) > 0.98


def test_normalize_cast_strided_and_unaligned(eng):
    import torch
    x = torch.randn(50, 200, device="cuda")
    view = x[:, 3:103]                                # ld=200, offset 3 floats -> scalar path
    out = eng.normalize_cast(view)
    torch.cuda.synchronize()
    want = O.bf16_round(O.l2_normalize(view.cpu().numpy()))
    got = out.float().cpu().numpy()[:, :100]
    assert np.max(np.abs(got - want)) <= 2.0 ** -7


@pytest.mark.parametrize("d", [100, 512, 4096, 8448, 49152])
@pytest.mark.parametrize("half", ["float16", "bfloat16"])
def test_normalize_cast_half_input_equals_its_fp32_image(eng, half, d):
    """fp16 / bf16 descriptors (an extractor under autocast, place_recognition.py:291-297) are widened exactly and
    summed in the fp32 form's order: bit-identical rows, both K1 forms (one block per row, cluster per row)."""
    import torch
    n = 37 if d > 8448 else 301
    g = torch.Generator(device="cuda").manual_seed(d)
    x = (torch.randn(n, d, device="cuda", generator=g) * (torch.rand(n, 1, device="cuda", generator=g) * 30 + 0.1)).to(getattr(torch, half))
    x[1] = 0
    got = eng.normalize_cast(x)
    want = eng.normalize_cast(x.float())
    torch.cuda.synchronize()
    assert got.dtype == torch.bfloat16 and torch.equal(got.view(torch.int16), want.view(torch.int16))
    # strided rows, unaligned start: the scalar path
    v = x[:, 3:3 + min(d - 3, 1001)]
    assert torch.equal(eng.normalize_cast(v).view(torch.int16), eng.normalize_cast(v.float().contiguous()).view(torch.int16))
    with pytest.raises(TypeError):
        eng.normalize_cast(x.double())


def test_host_abi_half_precision_descriptors(eng, monkeypatch):
    """semgate_find_loop_closures_host_dtype: fp16 (numpy) and bf16 (CPU torch tensor) host databases give the fp32
    call's candidates on the widened rows, bit for bit, chunked pipeline included; and the oracle's on those rows."""
    import torch
    from semgate import _native, synthetic
    desc, ts, fl = synthetic.make_case(9000, 2048, 3, seed=46)          # 74 MB of fp32 -> 3 row chunks
    fl32 = fl.astype(np.int32)
    p = _native.make_params(k=25, similarity_threshold=0.5, min_time_gap=10.0, max_floor_diff=0)
    d16 = desc.astype(np.float16)
    want = eng.find_loop_closures_host(d16.astype(np.float32), ts, fl32, p)
    got = eng.find_loop_closures_host(d16, ts, fl32, p)
    assert len(want[0]) > 1000
    for a, b in zip(got, want):
        assert np.array_equal(a, b)
    monkeypatch.setenv("SEMGATE_E2E_CHUNKS", "3,1,2,1")                 # other row cuts: same lists
    for a, b in zip(eng.find_loop_closures_host(d16, ts, fl32, p), want):
        assert np.array_equal(a, b)
    monkeypatch.delenv("SEMGATE_E2E_CHUNKS")
    ref = O.find_loop_closures(d16.astype(np.float32), ts, fl32, similarity_threshold=0.5, min_time_gap=10.0, k=25, bf16=True)
    g = dict(query_idx=got[0].astype(np.int64), match_idx=got[1].astype(np.int64), similarity=got[2], is_valid=got[3])
    parity.check_order(g)
    rep = parity.compare_candidates(ref, g, 25, 0.5, tol=BF16_MODEL_TOL)
    assert rep["boundary_diffs"] <= 2
    parity.check_decisions_exact(g, ts, fl32, 10.0, 0)
    # bf16 rows in host memory
    tb = torch.from_numpy(desc[:3000]).to(torch.bfloat16)
    got = eng.find_loop_closures_host(tb, ts[:3000], fl32[:3000], p)
    want = eng.find_loop_closures_host(tb.float().numpy(), ts[:3000], fl32[:3000], p)
    assert len(want[0]) > 100
    for a, b in zip(got, want):
        assert np.array_equal(a, b)
    # the C ABI refuses element types it does not know
    tot = _native.C.c_int64(0)
    rc = eng.lib.semgate_find_loop_closures_host_dtype(eng._h, None, 7, 10, 64, None, None, _native.C.byref(p), None, None, None,
                                                       None, 0, _native.C.byref(tot))
    assert rc == _native.EINVAL


def test_mirrored_class_keeps_fp16_descriptors_fp16(eng):
    """PlaceDescriptor.descriptor arrays of dtype float16 are uploaded as they are (half the PCIe bytes) and give the
    lists of their fp32 image."""
    from semgate import SemanticPlaceRecognition, synthetic
    desc, ts, fl = synthetic.make_case(700, 512, 3, seed=47)
    d16 = desc.astype(np.float16)
    out = []
    for rows in (d16, d16.astype(np.float32)):
        spr = SemanticPlaceRecognition('mixvpr', 'cuda', similarity_threshold=0.5, min_time_gap=10.0, descriptor_dim=512)
        for i in range(len(rows)):
            spr.add_image(rows[i], float(ts[i]), int(fl[i]))
        out.append([(m.query_idx, m.match_idx, m.similarity, m.is_valid) for m in spr.find_loop_closures(k=10)])
    assert len(out[0]) > 100 and out[0] == out[1]


# --------------------------------------------------------------------------- K2/K3 vs oracle
SHAPES = [
    # Q,   N,    D,   k,  thr,  gap, floors
    (128,  256,  64,  5,  0.3,  2.0, 3),      # exactly one tile
    (100,  200,  64,  10, 0.5,  10.0, 3),     # ragged single tile
    (300,  700,  128, 25, 0.5,  10.0, 3),     # several tiles, ragged edges
    (257,  513,  96,  7,  0.2,  3.0, 4),      # D not a multiple of 64
    (1000, 1000, 512, 25, 0.5,  10.0, 3),     # C1 slice
    (64,   5000, 256, 32, 0.4,  5.0, 4),      # few queries, long database (split-N)
    (1500, 300,  64,  3,  0.5,  1.0, 3),      # many query blocks, short database
    (400,  2100, 320, 64, 0.1,  10.0, 3),     # k = 64
]


@pytest.mark.parametrize("cg", [1, 2, 4])
@pytest.mark.parametrize("shape", SHAPES, ids=[f"{s[0]}x{s[1]}x{s[2]}k{s[3]}" for s in SHAPES])
def test_gated_topk_vs_oracle(eng, shape, cg):
    from semgate import synthetic
    Q, N, D, k, thr, gap, nf = shape
    n = max(Q, N)
    desc, ts, fl = synthetic.make_case(n, D, nf, seed=Q + N + D)
    q, db = desc[:Q], desc[:N]
    fl32 = fl.astype(np.int32)
    got = run_gpu(eng, q, db, k, thr, gap, ts[:Q], ts[:N], fl32[:Q], fl32[:N], mfd=0, mode=0, cg=cg)
    check_padded(got, k)
    # (1) against the GPU arithmetic model: bf16 operands, wide accumulate -> tight
    ref16 = O.gated_topk(q, db, ts[:Q], ts[:N], fl32[:Q], fl32[:N], k=k, threshold=thr, min_time_gap=gap,
                         max_floor_diff=0, bf16=True)
    rep = parity.compare_candidates(O.compact(ref16), O.compact(got), k, thr, tol=BF16_MODEL_TOL)
    assert rep["max_score_err"] < BF16_MODEL_TOL
    # (2) against the reference arithmetic (fp32 operands): the north-star rule
    ref32 = O.gated_topk(q, db, ts[:Q], ts[:N], fl32[:Q], fl32[:N], k=k, threshold=thr, min_time_gap=gap,
                         max_floor_diff=0)
    parity.compare_candidates(O.compact(ref32), O.compact(got), k, thr, tol=parity.SCORE_TOL)
    # (3) bit-exact decisions on everything returned
    c = O.compact(got)
    parity.check_decisions_exact(c, ts[:N], fl32[:N], gap, 0, q_ts=ts[:Q], q_floors=fl32[:Q])


@pytest.mark.parametrize("seed", list(range(16)))
def test_randomised_configurations_vs_oracle(eng, seed):
    """Seeded random draws over everything the interface lets a caller vary (shapes incl. ragged and tiny,
    k, threshold incl. none, window incl. 0 / none, strict / non-strict / no gate, flag / mask mode, None
    labels, unsorted timestamps, tile shape 1 / 2 / 4 CTAs): exact candidate sets against the oracle's
    model of the GPU arithmetic, bit-exact decisions on every returned pair."""
    from semgate import synthetic
    rng = np.random.default_rng(1000 + seed)
    Q = int(rng.choice([1, 3, 127, 129, 300, 641, 1100]))
    N = int(rng.choice([1, 2, 255, 257, 900, 2500, 6000]))
    D = int(rng.choice([17, 64, 100, 192, 500, 1031]))
    k = int(rng.choice([1, 5, 10, 25, 33, 64]))
    thr = float(rng.choice([-np.inf, 0.2, 0.45, 0.6]))
    gap = float(rng.choice([0.0, 2.5, 10.0]))
    use_ts = bool(rng.integers(0, 4))           # 1 in 4: no timestamps at all (query(timestamp=None))
    mfd = int(rng.choice([-1, 0, 1]))
    mode = int(rng.integers(0, 2))
    cg = int(rng.choice([1, 2, 4]))
    n = max(Q, N)
    desc, ts, fl = synthetic.make_case(n, D, int(rng.integers(2, 6)), seed=seed)
    fl32 = fl.astype(np.int32)
    fl32[rng.uniform(size=n) < 0.05] = O.FLOOR_NONE
    if rng.integers(0, 2):
        ts = ts[rng.permutation(n)]             # arbitrary (unsorted) stamps
    qts, dts = (ts[:Q], ts[:N]) if use_ts else (None, None)
    got = run_gpu(eng, desc[:Q], desc[:N], k, thr, gap, qts, dts, fl32[:Q], fl32[:N], mfd=mfd, mode=mode, cg=cg)
    check_padded(got, k)
    ref = O.gated_topk(desc[:Q], desc[:N], qts, dts, fl32[:Q], fl32[:N], k=k, threshold=thr, min_time_gap=gap,
                       max_floor_diff=mfd, gate_mode=O.GATE_MASK if mode else O.GATE_FLAG, bf16=True)
    # one bf16 rounding flip of a normalised element (row norm summed in a different order) moves a score by up
    # to ~ 2^-9 * x_i * y_i, which grows as the descriptor gets shorter: 4.9e-4 at d = 17
    parity.compare_candidates(O.compact(ref), O.compact(got), k, thr, tol=max(BF16_MODEL_TOL, 8e-3 / D))
    c = O.compact(got)
    if len(c["query_idx"]):
        if use_ts:
            parity.check_decisions_exact(c, ts[:N], fl32[:N], gap, mfd, q_ts=ts[:Q], q_floors=fl32[:Q])
        if mode == 1 and mfd >= 0:
            assert c["is_valid"].all()


@pytest.mark.parametrize("Q", [1, 2, 3, 4])
def test_streaming_query_kernel(eng, Q):
    """K6 (<= 4 query rows, cta_group auto): the bandwidth-built GEMV form must give what the fused
    tensor-core kernel and the oracle give — every gate / window / threshold / top-k rule, ragged row
    splits, shard offsets, k up to 64."""
    import torch
    from semgate import _native, synthetic
    for N, D, k, thr, gap, mode, mfd in ((1, 64, 5, -np.inf, 10.0, 0, 0), (37, 100, 25, 0.2, 0.0, 0, -1),
                                         (5000, 512, 25, 0.45, 10.0, 0, 0), (20011, 4096, 64, 0.3, 3.0, 1, 1),
                                         (3000, 8448, 10, 0.4, 10.0, 1, 0)):
        n = max(N, Q)
        desc, ts, fl = synthetic.make_case(n, D, 4, seed=N + Q)
        fl32 = fl.astype(np.int32)
        fl32[::11] = O.FLOOR_NONE
        # queries: perturbed copies of database rows spread over the database (so every block range has hits)
        qrows = (np.arange(Q) * max(1, N // max(Q, 1)) + N // 3) % N
        qd = desc[qrows] + 0.05 * np.random.default_rng(Q).standard_normal((Q, D)).astype(np.float32)
        qts = ts[qrows] + 1.0
        qfl = fl32[qrows]
        args = (qd, desc[:N], k, thr, gap, qts, ts[:N], qfl, fl32[:N])
        got = run_gpu(eng, *args, mfd=mfd, mode=mode, cg=0, offset=7)          # auto -> K6
        check_padded(got, k)
        k2 = run_gpu(eng, *args, mfd=mfd, mode=mode, cg=1, offset=7)           # pinned tile shape -> K2
        ref = O.gated_topk(qd, desc[:N], qts, ts[:N], qfl, fl32[:N], k=k, threshold=thr, min_time_gap=gap,
                           max_floor_diff=mfd, gate_mode=O.GATE_MASK if mode else O.GATE_FLAG, bf16=True, db_index_offset=7)
        tol = max(BF16_MODEL_TOL, 8e-3 / D)
        parity.compare_candidates(O.compact(ref), O.compact(got), k, thr, tol=tol)
        parity.compare_candidates(O.compact(k2), O.compact(got), k, thr, tol=5e-5)   # same operands; CUDA-core fp32 FMA chain vs tensor-core accumulation
        assert np.array_equal(got["keys"], O.pack_keys(got["scores"], got["idx"]))


@pytest.mark.parametrize("cg", [1, 2, 4])
def test_mask_mode_and_nonstrict(eng, cg):
    from semgate import synthetic
    desc, ts, fl = synthetic.make_case(900, 128, 4, seed=77)
    fl32 = fl.astype(np.int32)
    for mfd in (0, 1):
        got = run_gpu(eng, desc, desc, 6, 0.3, 4.0, ts, ts, fl32, fl32, mfd=mfd, mode=1, cg=cg)
        check_padded(got, 6)
        ref = O.gated_topk(desc, desc, ts, ts, fl32, fl32, k=6, threshold=0.3, min_time_gap=4.0,
                           max_floor_diff=mfd, gate_mode=O.GATE_MASK, bf16=True)
        parity.compare_candidates(O.compact(ref), O.compact(got), 6, 0.3, tol=BF16_MODEL_TOL)
        c = O.compact(got)
        assert c["is_valid"].all(), "mask mode returns floor-consistent pairs only"
        parity.check_decisions_exact(c, ts, fl32, 4.0, mfd)


def test_none_floors_and_gating_off(eng):
    from semgate import synthetic
    desc, ts, fl = synthetic.make_case(500, 64, 3, seed=5)
    fl32 = fl.astype(np.int32)
    fl32[::7] = O.FLOOR_NONE
    got = run_gpu(eng, desc, desc, 8, 0.4, 4.0, ts, ts, fl32, fl32, mfd=0)
    c = O.compact(got)
    parity.check_decisions_exact(c, ts, fl32, 4.0, 0)
    assert c["is_valid"][(fl32[c["query_idx"]] == O.FLOOR_NONE)].all()
    off = run_gpu(eng, desc, desc, 8, 0.4, 4.0, ts, ts, fl32, fl32, mfd=-1)
    assert O.compact(off)["is_valid"].all()
    nofl = run_gpu(eng, desc, desc, 8, 0.4, 4.0, ts, ts, None, None, mfd=0)
    assert O.compact(nofl)["is_valid"].all()
    assert np.array_equal(nofl["idx"], got["idx"])


def test_no_timestamps_and_unsorted_timestamps(eng):
    from semgate import synthetic
    desc, ts, fl = synthetic.make_case(600, 64, 3, seed=9)
    got = run_gpu(eng, desc[:50], desc, 5)                      # query(timestamp=None): self match on top
    assert np.array_equal(got["idx"][:, 0], np.arange(50))
    assert np.allclose(got["scores"][:, 0], 1.0, atol=4e-3)
    perm = np.random.default_rng(0).permutation(600)           # unsorted time stamps: per-pair fp64 predicate
    tsp = ts[perm]
    got = run_gpu(eng, desc, desc, 9, 0.3, 6.0, tsp, tsp)
    ref = O.gated_topk(desc, desc, tsp, tsp, k=9, threshold=0.3, min_time_gap=6.0, max_floor_diff=-1, bf16=True)
    parity.compare_candidates(O.compact(ref), O.compact(got), 9, 0.3, tol=BF16_MODEL_TOL)
    parity.check_decisions_exact(O.compact(got), tsp, None, 6.0, -1)


def test_epoch_scale_timestamps_need_fp64(eng):
    """fp32 cannot resolve the window at epoch scale (128 s ulp); the decisions must
    follow the fp64 predicate exactly, including pairs exactly `gap` apart (kept: strict <)."""
    from semgate import synthetic
    desc = synthetic.make_descriptors(400, 64, seed=3, places=4)     # many high-similarity pairs
    ts = synthetic.EPOCH0 + 0.25 * np.arange(400)
    got = run_gpu(eng, desc, desc, 25, -1.0, 2.5, ts, ts)
    ref = O.gated_topk(desc, desc, ts, ts, k=25, threshold=-1.0, min_time_gap=2.5, max_floor_diff=-1, bf16=True)
    parity.compare_candidates(O.compact(ref), O.compact(got), 25, -1.0, tol=BF16_MODEL_TOL)
    c = O.compact(got)
    parity.check_decisions_exact(c, ts, None, 2.5, -1)
    d = np.abs(ts[c["match_idx"]] - ts[c["query_idx"]])
    assert d.min() >= 2.5
    ro = O.compact(ref)
    dr = np.abs(ts[ro["match_idx"]] - ts[ro["query_idx"]])
    assert (d == 2.5).sum() == (dr == 2.5).sum(), "pairs exactly min_time_gap apart are kept (strict <)"


@pytest.mark.parametrize("D,cg,sym", [(64, 1, 0), (64, 2, 0), (192, 2, 0), (2048, 2, 0), (1024, 2, 1)])
def test_window_chunks_of_a_real_sequence(eng, D, cg, sym):
    """A sequence whose temporal neighbours look alike (a slow random walk of the descriptor, 10 Hz keyframes, 10 s
    window): every row meets several 32-column chunks that lie wholly inside its exclusion window and full of scores above
    the threshold -- the case the epilogue's chunk-level window test skips.  With stamps that break the pattern (NaN,
    +-inf, a block out of order, a run of equal stamps, a window edge on a chunk border) the lists must equal the
    oracle's and every decision the fp64 predicate, for one and two epilogue sets, CTA pairs and the symmetric sweep."""
    n, k, gap, thr = 1500, 25, 10.0, 0.6
    rng = np.random.default_rng(D + cg)
    steps = rng.standard_normal((n, D)).astype(np.float32)
    desc = np.cumsum(0.05 * steps, axis=0) + rng.standard_normal((1, D)).astype(np.float32)     # neighbours: cosine ~ 0.99
    desc[700:] += 3.0 * rng.standard_normal((1, D)).astype(np.float32)                         # a second place
    desc[1100:1300] = desc[300:500] + 0.02 * rng.standard_normal((200, D)).astype(np.float32)  # a revisit: real loop closures
    ts = 1000.0 + 0.1 * np.arange(n)
    ts[64] = np.nan
    ts[200] = np.inf
    ts[333] = -np.inf
    ts[400:464] = ts[400:464][::-1].copy()        # out of order inside a block
    ts[600:640] = ts[600]                        # equal stamps
    ts[900] = ts[800] + gap                       # exactly `gap` from row 800: kept (strict <)
    got = run_gpu(eng, desc, desc, k, thr, gap, ts, ts, cg=cg, sym=sym)
    assert got["mode"] == sym
    ref = O.gated_topk(desc, desc, ts, ts, k=k, threshold=thr, min_time_gap=gap, max_floor_diff=-1, bf16=True)
    c = O.compact(got)
    assert len(c["query_idx"]) > 2000
    parity.compare_candidates(O.compact(ref), c, k, thr, tol=BF16_MODEL_TOL)
    parity.check_decisions_exact(c, ts, None, gap, -1)
    # the same sweep without the window: the neighbours come back, i.e. the window did decide something
    free = O.compact(run_gpu(eng, desc, desc, k, thr, 0.0, ts, ts, cg=cg, sym=sym))
    near = lambda cc: int((np.abs(cc["match_idx"] - cc["query_idx"]) < 50).sum())
    assert near(free) > 10000 and near(c) < near(free) // 20


@pytest.mark.parametrize("cg", [1, 2])
def test_two_epilogue_sets_give_the_same_lists(eng, monkeypatch, cg):
    """SEMGATE_EPI_SETS=2 (opt-in): two epilogue warps per TMEM quarter, the two threads of a row share its list under a
    lock.  Same keys as one set, bit for bit: sparse hits, a dense sequence (every row hundreds of hits: both threads
    insert all the time), mask mode, k = 64, a ragged rectangular sweep, a multi-pass k = 100 sweep."""
    from semgate import synthetic
    rng = np.random.default_rng(cg)
    n, D = 1300, 96
    desc, ts, fl = synthetic.make_case(n, D, 3, seed=5)
    fl32 = fl.astype(np.int32)
    dense = np.repeat(rng.standard_normal((13, D)).astype(np.float32), 100, axis=0) + 0.2 * rng.standard_normal((n, D)).astype(np.float32)
    cases = [dict(q=desc, db=desc, k=25, thr=0.3, gap=5.0, mfd=0, mode=0), dict(q=dense, db=dense, k=25, thr=0.5, gap=2.0, mfd=-1, mode=0),
             dict(q=dense, db=dense, k=64, thr=-1.0, gap=0.0, mfd=0, mode=1), dict(q=desc[:333], db=desc[:1111], k=9, thr=0.2, gap=5.0, mfd=1, mode=0),
             dict(q=dense[:200], db=dense, k=100, thr=0.4, gap=1.0, mfd=-1, mode=0)]
    for c in cases:
        nq, nd = len(c["q"]), len(c["db"])
        outs = {}
        for sets in ("1", "2"):
            monkeypatch.setenv("SEMGATE_EPI_SETS", sets)
            outs[sets] = run_gpu(eng, c["q"], c["db"] if c["db"] is not c["q"] else c["q"], c["k"], c["thr"], c["gap"], ts[:nq], ts[:nd] if nd != nq else ts[:nq],
                                 fl32[:nq], fl32[:nd] if nd != nq else fl32[:nq], mfd=c["mfd"], mode=c["mode"], cg=cg, sym=-1)
        monkeypatch.delenv("SEMGATE_EPI_SETS")
        assert np.array_equal(outs["1"]["keys"], outs["2"]["keys"]) and np.array_equal(outs["1"]["valid"], outs["2"]["valid"])
        assert outs["1"]["count"].sum() > 500


def test_adversarial_ascending_scores(eng):
    """Every new column beats the current k-th score (threshold -1, similarity strictly rising
    with the index): the running list is replaced on every element.  Operands are fed as raw
    bf16 so that all 3000 scores are exactly representable and distinct."""
    import torch
    from semgate import _native
    D, N, k = 64, 3000, 25
    i = np.arange(N)
    db = np.zeros((N, D), np.float32)
    db[:, 0] = (128 + i // 32) / 256.0            # bf16-exact
    db[:, 1] = (i % 32) / 32.0                    # bf16-exact
    q = np.zeros((130, D), np.float32)
    q[:, 0] = 1.0
    q[:, 1] = 1.0 / 256.0                         # score_i = (128 + i//32)/256 + (i%32)/8192, exact in fp32
    qb = torch.from_numpy(q).cuda().to(torch.bfloat16)
    dbb = torch.from_numpy(db).cuda().to(torch.bfloat16)
    assert torch.equal(dbb.float().cpu(), torch.from_numpy(db))
    for cg in (1, 2):
        r = eng.gated_topk(qb, dbb, _native.make_params(k=k, similarity_threshold=-1.0, cta_group=cg))
        torch.cuda.synchronize()
        got = dict(scores=r.scores.cpu().numpy(), idx=r.idx.cpu().numpy().astype(np.int64),
                   valid=r.valid.cpu().numpy().astype(bool), count=r.count.cpu().numpy())
        check_padded(got, k)
        want = np.arange(N - 1, N - 1 - k, -1)
        assert np.array_equal(got["idx"][0], want) and np.array_equal(got["idx"][129], want)
        assert np.array_equal(got["scores"][0], (db[want, 0] + db[want, 1] / 256.0).astype(np.float32))


def test_ties_prefer_lower_index(eng):
    rng = np.random.default_rng(1)
    v = rng.standard_normal((1, 64)).astype(np.float32)
    db = np.repeat(v, 700, axis=0)                      # identical rows -> identical scores
    got = run_gpu(eng, v, db, 10)
    assert np.array_equal(got["idx"][0], np.arange(10))


def test_k_larger_than_database_and_tiny_inputs(eng):
    from semgate import synthetic
    desc, ts, fl = synthetic.make_case(6, 32, 3, seed=2)
    got = run_gpu(eng, desc, desc, 25, -1.0, 0.6, ts, ts)
    assert got["count"].max() <= 5
    ref = O.gated_topk(desc, desc, ts, ts, k=25, threshold=-1.0, min_time_gap=0.6, max_floor_diff=-1, bf16=True)
    parity.compare_candidates(O.compact(ref), O.compact(got), 25, -1.0, tol=BF16_MODEL_TOL)
    one = run_gpu(eng, desc[:1], desc[:1], 3)
    assert one["count"][0] == 1 and one["idx"][0, 0] == 0


def test_db_index_offset_and_merge(eng):
    """Row-sharded database: two shard sweeps + key merge == one sweep."""
    import torch
    from semgate import synthetic
    desc, ts, fl = synthetic.make_case(1100, 128, 3, seed=31)
    fl32 = fl.astype(np.int32)
    whole = run_gpu(eng, desc, desc, 12, 0.3, 5.0, ts, ts, fl32, fl32, mfd=0)
    cut = 600
    a = run_gpu(eng, desc, desc[:cut], 12, 0.3, 5.0, ts, ts[:cut], fl32, fl32[:cut], mfd=0, offset=0)
    b = run_gpu(eng, desc, desc[cut:], 12, 0.3, 5.0, ts, ts[cut:], fl32, fl32[cut:], mfd=0, offset=cut)
    assert np.array_equal(a["keys"], O.pack_keys(a["scores"], a["idx"])), "key wire format differs from the oracle's"
    keys = torch.from_numpy(np.stack([a["keys"], b["keys"]])).cuda()
    m = eng.merge_topk(keys, 12, q_floor=_t(fl32), db_floor_all=_t(fl32), max_floor_diff=0)
    torch.cuda.synchronize()
    assert np.array_equal(m.idx.cpu().numpy(), whole["idx"])
    assert np.array_equal(m.scores.cpu().numpy(), whole["scores"])
    assert np.array_equal(m.valid.cpu().numpy().astype(bool), whole["valid"])
    assert np.array_equal(m.count.cpu().numpy(), whole["count"])


@pytest.mark.parametrize("G,k,fill", [(1, 25, 1.0), (4, 25, 0.6), (5, 25, 1.0), (6, 25, 1.0), (8, 64, 0.9), (29, 25, 0.3), (3, 1, 1.0)])
def test_merge_topk_kernel_exact(eng, G, k, fill):
    """K3 against a numpy sort of the same keys: one batch (<= 128 keys, 4 per lane), multi-batch
    (8 per lane), partially filled and empty lists, unsorted inputs."""
    import torch
    rng = np.random.default_rng(G * 100 + k)
    Q = 777
    sc = rng.uniform(-1, 1, size=(G, Q, k)).astype(np.float32)
    ix = rng.permutation(G * Q * k).reshape(G, Q, k).astype(np.int64) % (2 ** 31 - 1)
    ix[rng.uniform(size=ix.shape) > fill] = -1                  # empty slots anywhere in a list
    ix[:, 5] = -1                                               # a row with no candidates at all
    keys = O.pack_keys(sc, ix)
    m = eng.merge_topk(torch.from_numpy(keys).cuda(), k, want_keys=True)
    torch.cuda.synchronize()
    flat = keys.view(np.uint64).transpose(1, 0, 2).reshape(Q, G * k)
    want = np.sort(flat, axis=1)[:, ::-1][:, :k]                # uint64 sort, descending
    got = m.keys.cpu().numpy().view(np.uint64)
    pad = np.zeros((Q, max(0, k - want.shape[1])), np.uint64)
    assert np.array_equal(got, np.concatenate([want, pad], axis=1))
    cnt = (want != 0).sum(axis=1)
    assert np.array_equal(m.count.cpu().numpy(), cnt)
    assert m.count.cpu().numpy()[5] == 0 and np.all(m.idx.cpu().numpy()[5] == -1)


@pytest.mark.parametrize("G,k,order", [(4, 25, "sorted"), (8, 25, "sorted"), (4, 25, "mixed"), (3, 32, "unsorted"), (2, 7, "sorted")])
def test_merge_many_rows_sorted_lists(eng, G, k, order):
    """K3's thread-per-row network kernel (many rows, a few lists) skips the sorting network for lists that arrive sorted
    (the per-GPU lists of a sharded sweep are K3 outputs): sorted, unsorted and mixed inputs (some warps take the short
    cut, some do not; short and empty lists) against a numpy sort of the same keys."""
    import torch
    rng = np.random.default_rng(G * 1000 + k)
    Q = 65536 + 77
    keys = rng.integers(1, 2 ** 62, size=(G, Q, k), dtype=np.int64).view(np.uint64)
    fill = rng.integers(0, k + 1, size=(G, Q))
    fill[:, ::5] = k                                             # plenty of full lists
    keys[np.arange(k)[None, None, :] >= fill[:, :, None]] = 0   # lists end early (0 = empty), row 3 of list 0 is empty
    keys[0, 3] = 0
    if order in ("sorted", "mixed"):
        keys = np.sort(keys, axis=2)[:, :, ::-1].copy()
    if order == "mixed":
        sel = rng.uniform(size=Q) < 0.02                        # a few rows per warp break the order
        shuf = keys[:, sel].copy()
        rng.shuffle(shuf, axis=2)
        keys[:, sel] = shuf
    m = eng.merge_topk(torch.from_numpy(keys.view(np.int64)).cuda(), k, want_keys=True)
    torch.cuda.synchronize()
    flat = keys.transpose(1, 0, 2).reshape(Q, G * k)
    want = np.sort(flat, axis=1)[:, ::-1][:, :k]
    got = m.keys.cpu().numpy().view(np.uint64)
    assert np.array_equal(got, want)
    assert np.array_equal(m.count.cpu().numpy(), (want != 0).sum(axis=1))


@pytest.mark.parametrize("G,k,Q,order", [(4, 25, 65536 + 77, "sorted"), (8, 25, 65536 + 31, "mixed"), (3, 32, 65536 + 1, "sorted"),
                                         (2, 10, 70000, "sorted"), (5, 7, 65537, "unsorted"), (1, 25, 65536, "sorted"),
                                         (4, 25, 65536 + 96, "sparse"), (3, 5, 65536 + 5, "mixed"), (8, 10, 65536 + 130, "unsorted"),
                                         (2, 1, 65536, "sorted"), (3, 30, 65536 + 64, "sorted")])
def test_merge_dense_lists_kernel(eng, G, k, Q, order):
    """K3's dense-list kernel (lists handed over as arrays: one bulk copy per warp and list, no packing pass, sorted lists
    folded without the sorting network): odd k (bulk copies), even k (pitched rows), k = 32, ragged last warps, one list,
    unsorted and nearly empty lists, a seeded list, a row slice that starts at an odd row (unaligned sources), the pointer
    table form, strided outputs of a k > 64 pass -- against a numpy sort and against the general network kernel."""
    import torch
    rng = np.random.default_rng(G * 1000 + k + Q)
    keys = rng.integers(1, 2 ** 62, size=(G, Q, k), dtype=np.int64).view(np.uint64)
    fill = rng.integers(0, k + 1, size=(G, Q))
    if order == "sparse":
        fill = rng.integers(0, 3, size=(G, Q))
    else:
        fill[:, ::3] = k
    keys[np.arange(k)[None, None, :] >= fill[:, :, None]] = 0
    keys[0, 3] = 0
    if order != "unsorted":
        keys = np.sort(keys, axis=2)[:, :, ::-1].copy()
    if order == "mixed":
        sel = rng.uniform(size=Q) < 0.02
        shuf = keys[:, sel].copy()
        rng.shuffle(shuf, axis=2)
        keys[:, sel] = shuf
    tk = torch.from_numpy(keys.view(np.int64)).cuda()
    nfl = 2 ** 20
    fl = torch.from_numpy(rng.integers(1, 4, size=nfl).astype(np.int32)).cuda()
    qf = torch.from_numpy(rng.integers(1, 4, size=Q).astype(np.int32)).cuda()
    keys_m = keys.copy()
    keys_m[keys_m != 0] = (keys_m[keys_m != 0] & ~np.uint64(0xFFFFFFFF)) | (np.uint64(0xFFFFFFFF) - (keys_m[keys_m != 0] % np.uint64(nfl)))
    # ^ indices inside the label array (key = score bits << 32 | ~index); the order inside a list may change: re-sort
    if order not in ("unsorted",):
        keys_m = np.sort(keys_m, axis=2)[:, :, ::-1].copy()
    tkm = torch.from_numpy(keys_m.view(np.int64)).cuda()
    m = eng.merge_topk(tkm, k, q_floor=qf, db_floor_all=fl, max_floor_diff=0, want_keys=True)
    torch.cuda.synchronize()
    flat = keys_m.transpose(1, 0, 2).reshape(Q, G * k)
    want = np.sort(flat, axis=1)[:, ::-1][:, :k]
    got = m.keys.cpu().numpy().view(np.uint64)
    assert np.array_equal(got, want)
    assert np.array_equal(m.count.cpu().numpy(), (want != 0).sum(axis=1))
    widx = np.where(want != 0, (np.uint64(0xFFFFFFFF) - (want & np.uint64(0xFFFFFFFF))).astype(np.int64), -1)
    assert np.array_equal(m.idx.cpu().numpy(), widx)
    wvalid = (want != 0) & (fl.cpu().numpy()[np.maximum(widx, 0)] == qf.cpu().numpy()[:, None])
    assert np.array_equal(m.valid.cpu().numpy().astype(bool), wvalid)
    # the general network kernel on the same lists: every output array bit for bit
    eng.set_option("k3_dense", 0)
    try:
        n = eng.merge_topk(tkm, k, q_floor=qf, db_floor_all=fl, max_floor_diff=0, want_keys=True)
        torch.cuda.synchronize()
    finally:
        eng.set_option("k3_dense", 1)
    def same(x, y, f):       # random keys decode to NaN scores now and then: compare bits
        x, y = getattr(x, f), getattr(y, f)
        return torch.equal(x.view(torch.int32), y.view(torch.int32)) if f == "scores" else torch.equal(x, y)
    for f in ("keys", "scores", "idx", "valid", "count"):
        assert same(m, n, f), f
    # the pointer-table form on a row slice that starts at an odd row (sources 8 bytes off a 16-byte boundary for odd k)
    parts = [tkm[g].clone() for g in range(G)]
    table = torch.tensor([p.data_ptr() for p in parts], dtype=torch.int64, device="cuda")
    lo, rows = 33, Q - 40
    mine = eng.merge_topk_peers_rows(table.data_ptr(), G, Q, k, lo, rows, q_floor=qf, db_floor_all=fl, max_floor_diff=0, want_keys=True)
    torch.cuda.synchronize()
    for f in ("keys", "idx", "valid", "count"):
        assert torch.equal(getattr(mine, f), getattr(m, f)[lo:lo + rows]), f
    assert torch.equal(mine.scores.view(torch.int32), m.scores[lo:lo + rows].view(torch.int32))
    even = eng.merge_topk_peers_rows(table.data_ptr(), G, Q, k, 64, Q - 64, q_floor=qf, db_floor_all=fl, max_floor_diff=0, want_keys=True)
    torch.cuda.synchronize()
    assert torch.equal(even.keys, m.keys[64:]) and torch.equal(even.valid, m.valid[64:])


def test_valid_only_compaction_and_device_statistics(eng):
    """§8f rank 4: the floor-consistent hand-off list (geometric_verification.py:709 skips cross-floor pairs) and
    get_statistics (place_recognition.py:913-933) reduced on the device, against the host path and the oracle."""
    from semgate import SemanticPlaceRecognition, PlaceDescriptor, synthetic
    n, d, k = 3000, 256, 25
    desc, ts, fl = synthetic.make_case(n, d, 3, seed=77)
    spr = SemanticPlaceRecognition('mixvpr', 'cuda', similarity_threshold=0.5, min_time_gap=10.0, descriptor_dim=d)
    for i in range(n):
        spr.vpr.descriptors.append(PlaceDescriptor(float(ts[i]), desc[i], floor_label=int(fl[i]) if i % 97 else None))
    full = spr.find_loop_closures_arrays(k=k)
    only, stats = spr.find_loop_closures_arrays(k=k, valid_only=True, with_statistics=True)
    keep = full.is_valid
    assert keep.sum() > 100 and (~keep).sum() > 100
    assert np.array_equal(only.query_idx, full.query_idx[keep]) and np.array_equal(only.match_idx, full.match_idx[keep])
    assert np.array_equal(only.similarity, full.similarity[keep]) and only.is_valid.all()
    host = spr.get_statistics(full)
    for key in ("total_matches", "valid_matches", "rejected_matches"):
        assert stats[key] == host[key]
    for key in ("rejection_rate", "mean_similarity", "mean_valid_similarity"):
        assert abs(stats[key] - host[key]) <= 1e-12 * max(1.0, abs(host[key]))
    full2, stats2 = spr.find_loop_closures_arrays(k=k, with_statistics=True)
    assert np.array_equal(full2.match_idx, full.match_idx) and stats2 == stats      # order-independent reduction
    # nothing valid / nothing at all
    spr2 = SemanticPlaceRecognition('mixvpr', 'cuda', similarity_threshold=2.0, descriptor_dim=d)
    for i in range(300):
        spr2.vpr.descriptors.append(PlaceDescriptor(float(ts[i]), desc[i], floor_label=int(fl[i])))
    e, st = spr2.find_loop_closures_arrays(k=k, valid_only=True, with_statistics=True)
    assert len(e) == 0 and st == {'total_matches': 0, 'valid_matches': 0, 'rejected_matches': 0, 'rejection_rate': 0.0}


def test_merge_topk_peers_pointer_table(eng):
    """The peer-memory form of K3 (lists addressed through a device pointer table, as over NVLink)
    equals the gathered form on the same keys; here all "peers" live on this GPU."""
    import torch
    rng = np.random.default_rng(5)
    for G, Q, k in ((2, 1000, 25), (8, 333, 25), (5, 64, 64), (148, 3, 25)):
        sc = rng.uniform(-1, 1, size=(G, Q, k)).astype(np.float32)
        ix = (rng.permutation(G * Q * k).reshape(G, Q, k) % (2 ** 31 - 1)).astype(np.int64)
        ix[rng.uniform(size=ix.shape) > 0.8] = -1
        keys = torch.from_numpy(O.pack_keys(sc, ix)).cuda()
        parts = [keys[g].clone() for g in range(G)]                      # separately allocated buffers
        table = torch.tensor([p.data_ptr() for p in parts], dtype=torch.int64, device="cuda")
        fl = torch.from_numpy(rng.integers(1, 5, size=2 ** 20).astype(np.int32)).cuda()
        qf = torch.from_numpy(rng.integers(1, 5, size=Q).astype(np.int32)).cuda()
        a = eng.merge_topk(keys, k, q_floor=qf, db_floor_all=None, max_floor_diff=-1, want_keys=True)
        b = eng.merge_topk_peers(table.data_ptr(), G, Q, k, q_floor=qf, db_floor_all=None, max_floor_diff=-1, want_keys=True)
        torch.cuda.synchronize()
        assert torch.equal(a.keys, b.keys) and torch.equal(a.idx, b.idx) and torch.equal(a.scores, b.scores)
        assert torch.equal(a.count, b.count) and torch.equal(a.valid, b.valid)
        del fl


def test_similarity_matrix_dense(eng):
    """Dense-output mode of the fused kernel vs numpy (place_recognition.py:171, :190), ragged shapes,
    and the mirrored _compute_similarity / compute_all_pairwise_similarities."""
    import torch
    from semgate import BasePlaceRecognition, PlaceDescriptor, synthetic
    for (Q, N, D) in ((300, 1000, 512), (1, 777, 100), (129, 257, 64), (5, 4099, 4096)):
        desc, _, _ = synthetic.make_case(max(Q, N), D, 3, seed=Q + N)
        qb = eng.normalize_cast(_t(desc[:Q]))
        db = eng.normalize_cast(_t(desc[:N]))
        got = eng.similarity_matrix(qb, db).cpu().numpy()
        a = qb.float().cpu().numpy().astype(np.float64)
        b = db.float().cpu().numpy().astype(np.float64)
        assert got.shape == (Q, N) and np.max(np.abs(got - a @ b.T)) <= 2e-5        # same bf16 operands
        ref = O.l2_normalize(desc[:Q]) @ O.l2_normalize(desc[:N]).T
        assert np.max(np.abs(got - ref)) <= parity.SCORE_TOL
    desc, ts, fl = synthetic.make_case(500, 256, 3, seed=3)
    vpr = BasePlaceRecognition(descriptor_dim=256)
    for i in range(500):
        vpr.descriptors.append(PlaceDescriptor(float(ts[i]), desc[i], floor_label=int(fl[i])))
    S = vpr.compute_all_pairwise_similarities()
    ref = O.l2_normalize(desc) @ O.l2_normalize(desc).T
    assert S.shape == (500, 500) and S.dtype == np.float32 and np.max(np.abs(S - ref)) <= parity.SCORE_TOL
    row = vpr._compute_similarity(desc[7] * 3.0, desc)                                # unnormalised query
    assert row.shape == (500,) and np.max(np.abs(row - ref[7])) <= parity.SCORE_TOL
    assert BasePlaceRecognition().compute_all_pairwise_similarities().size == 0


@pytest.mark.parametrize("cg", [1, 2, 4])
@pytest.mark.parametrize("chunks", ["1", "3"])
def test_paced_runs(eng, cg, chunks, monkeypatch):
    """The units of a super-row advance in lock-step through per-super-row chunk counters (L2 pacing).
    Forced on at test sizes with a window of 1 / 3 chunks: ragged tail super-rows, runs of different
    length sharing the counters, several chunks per tile (d = 2048 -> 2), one k-block per tile (d = 64).
    The result never depends on the pacing."""
    from semgate import synthetic
    cases = [(700, 45000, 64, 4), (9000, 9000, 64, 3), (1500, 5000, 2048, 3), (333, 7777, 192, 2)]
    for Q, N, D, F in cases:
        desc, ts, fl = synthetic.make_case(max(Q, N), D, F, seed=Q + D)
        fl32 = fl.astype(np.int32)
        args = (desc[:Q], desc[:N], 25, 0.45, 10.0, ts[:Q], ts[:N], fl32[:Q], fl32[:N])
        monkeypatch.setenv("SEMGATE_WINDOW_CHUNKS", chunks)
        got = run_gpu(eng, *args, mfd=0, cg=cg)
        monkeypatch.delenv("SEMGATE_WINDOW_CHUNKS")
        monkeypatch.setenv("SEMGATE_WINDOW_MB", "0")
        one = run_gpu(eng, *args, mfd=0, cg=cg)
        monkeypatch.delenv("SEMGATE_WINDOW_MB")
        check_padded(got, 25)
        assert np.array_equal(one["idx"], got["idx"]) and np.array_equal(one["scores"], got["scores"])
        if D == 64 and Q == 700:
            ref = O.gated_topk(desc[:Q], desc[:N], ts[:Q], ts[:N], fl32[:Q], fl32[:N], k=25, threshold=0.45,
                               min_time_gap=10.0, max_floor_diff=0, bf16=True)
            rep = parity.compare_candidates(O.compact(ref), O.compact(got), 25, 0.45, tol=BF16_MODEL_TOL)
            assert rep["boundary_diffs"] <= 2


def test_accumulate_over_database_slices(eng):
    """Sweeping disjoint database slices one after the other with `accumulate` == one sweep."""
    import torch
    from semgate import _native, synthetic
    desc, ts, fl = synthetic.make_case(2300, 128, 3, seed=8)
    fl32 = fl.astype(np.int32)
    xb = eng.normalize_cast(_t(desc))
    tts, tfl = _t(ts), _t(fl32)
    kw = dict(k=20, similarity_threshold=0.3, min_time_gap=5.0, max_floor_diff=0)
    whole = eng.gated_topk(xb, xb, _native.make_params(**kw), q_ts=tts, db_ts=tts, q_floor=tfl, db_floor=tfl, want_keys=True)
    keys = None
    for lo, hi in ((0, 700), (700, 1500), (1500, 2300)):
        p = _native.make_params(db_index_offset=lo, accumulate=keys is not None, **kw)
        # hand the previous keys back in: the binding allocates `keys` unless we pass them
        r = eng.gated_topk(xb, xb[lo:hi], p, q_ts=tts, db_ts=tts[lo:hi].contiguous(), q_floor=tfl,
                           db_floor=tfl[lo:hi].contiguous(), want_keys=True, want_lists=False, keys=keys)
        keys = r.keys
    torch.cuda.synchronize()
    assert torch.equal(keys, whole.keys)
    m = eng.merge_topk(keys.unsqueeze(0).contiguous(), 20, q_floor=tfl, db_floor_all=tfl, max_floor_diff=0)
    torch.cuda.synchronize()
    assert torch.equal(m.idx, whole.idx) and torch.equal(m.scores, whole.scores) and torch.equal(m.valid, whole.valid)


def test_accumulate_streaming_queries(eng):
    """`accumulate` through the streaming kernel (2 query rows) and the block-per-row merge with a seeded list."""
    import torch
    from semgate import _native, synthetic
    desc, ts, fl = synthetic.make_case(9000, 256, 3, seed=21)
    fl32 = fl.astype(np.int32)
    xb = eng.normalize_cast(_t(desc))
    tts, tfl = _t(ts), _t(fl32)
    q, qts, qfl = xb[100:102].contiguous(), tts[100:102].contiguous(), tfl[100:102].contiguous()
    kw = dict(k=25, similarity_threshold=0.2, min_time_gap=5.0, max_floor_diff=0)
    whole = eng.gated_topk(q, xb, _native.make_params(**kw), q_ts=qts, db_ts=tts, q_floor=qfl, db_floor=tfl, want_keys=True)
    keys = None
    for lo, hi in ((0, 4000), (4000, 4001), (4001, 9000)):
        p = _native.make_params(db_index_offset=lo, accumulate=keys is not None, **kw)
        r = eng.gated_topk(q, xb[lo:hi], p, q_ts=qts, db_ts=tts[lo:hi].contiguous(), q_floor=qfl,
                           db_floor=tfl[lo:hi].contiguous(), want_keys=True, want_lists=False, keys=keys)
        keys = r.keys
    torch.cuda.synchronize()
    assert int(whole.count.sum()) > 10
    assert torch.equal(keys, whole.keys)


def test_host_abi_chunked_pipeline(eng):
    """Large enough that semgate_find_loop_closures_host pipelines H2D chunks against partial sweeps."""
    from semgate import _native, synthetic
    desc, ts, fl = synthetic.make_case(9000, 2048, 3, seed=45)          # 74 MB -> 3 chunks
    fl32 = fl.astype(np.int32)
    p = _native.make_params(k=25, similarity_threshold=0.5, min_time_gap=10.0, max_floor_diff=0)
    q, m, s, v = eng.find_loop_closures_host(desc, ts, fl32, p)
    got = dict(query_idx=q.astype(np.int64), match_idx=m.astype(np.int64), similarity=s, is_valid=v)
    parity.check_order(got)
    ref = O.find_loop_closures(desc, ts, fl32, similarity_threshold=0.5, min_time_gap=10.0, k=25, bf16=True)
    rep = parity.compare_candidates(ref, got, 25, 0.5, tol=BF16_MODEL_TOL)
    assert rep["boundary_diffs"] <= 2
    parity.check_decisions_exact(got, ts, fl32, 10.0, 0)
    # and it equals the resident-data path exactly
    import torch
    xb = eng.normalize_cast(_t(desc))
    tts, tfl = _t(ts), _t(fl32)
    r = eng.gated_topk(xb, xb, p, q_ts=tts, db_ts=tts, q_floor=tfl, db_floor=tfl)
    oq, om, os_, ov, tot = eng.compact(r)
    t = int(tot.item())
    assert t == len(q) and np.array_equal(om[:t].cpu().numpy(), m) and np.array_equal(os_[:t].cpu().numpy(), s)


# --------------------------------------------------------------------------- reference goldens through the mirrored API
@pytest.mark.parametrize("path", FLC, ids=[os.path.basename(p)[:-4] for p in FLC])
def test_find_loop_closures_golden(path):
    from semgate import SemanticPlaceRecognition, PlaceDescriptor
    c = parity.load_flc_case(path)
    spr = SemanticPlaceRecognition('mixvpr', 'cuda', similarity_threshold=c["thr"], min_time_gap=c["gap"],
                                   descriptor_dim=c["d"])
    half = c["n"] // 2
    for i in range(half):                                # add_image path ...
        spr.add_image(c["desc"][i], float(c["ts"][i]), c["floors"][i])
    for i in range(half, c["n"]):                        # ... and direct appends, as the reference demo does
        spr.vpr.descriptors.append(PlaceDescriptor(float(c["ts"][i]), c["desc"][i], floor_label=c["floors"][i]))
    matches = spr.find_loop_closures(enable_floor_gating=c["gating"], k=c["k"])
    got = dict(query_idx=np.array([m.query_idx for m in matches], dtype=np.int64),
               match_idx=np.array([m.match_idx for m in matches], dtype=np.int64),
               similarity=np.array([m.similarity for m in matches]),
               is_valid=np.array([m.is_valid for m in matches], dtype=bool))
    parity.check_order(got)
    rep = parity.compare_candidates(c["ref"], got, c["k"], c["thr"])
    assert rep["max_score_err"] <= parity.SCORE_TOL
    enc = O.encode_floors(c["floors"])
    parity.check_decisions_exact(got, c["ts"], enc if c["gating"] else None, c["gap"], 0)
    assert all(m.query_timestamp == c["ts"][m.query_idx] and m.match_timestamp == c["ts"][m.match_idx] for m in matches[:200])
    st = spr.get_statistics(matches)
    assert st["total_matches"] == len(matches) and st["valid_matches"] == int(got["is_valid"].sum())
    # the composition the north star names: place recognition feeding the loop-closure gate
    if c["gating"] and not any(f is None for f in c["floors"]):
        from semgate import SemanticLoopClosureGate
        gate = SemanticLoopClosureGate(np.array(c["floors"]), strict_mode=True)
        valid, rejected = gate.gate_candidates([(m.query_idx, m.match_idx, m.similarity) for m in matches])
        assert len(valid) == st["valid_matches"] and len(rejected) == st["rejected_matches"]


@pytest.mark.parametrize("path", QRY, ids=[os.path.basename(p)[:-4] for p in QRY])
def test_query_golden(path):
    from semgate import BasePlaceRecognition, synthetic
    g = np.load(path)
    n, d, seed, k, with_ts = [int(v) for v in g["params"]]
    gap = float(g["fparams"][0])
    desc, ts, floors = synthetic.make_case(n + 4, d, 3, seed, 0.5)
    vpr = BasePlaceRecognition(descriptor_dim=d, device='cuda')
    for i in range(n):
        vpr.add_image(desc[i], float(ts[i]), int(floors[i]))
    for r, qi in enumerate(range(n, n + 4)):
        tq = float(ts[(qi * 97) % n]) + 0.25 if with_ts else None
        ms = vpr.query(desc[qi], tq, k=k, min_time_gap=gap)
        cnt = int(g["count"][r])
        ref = dict(query_idx=np.zeros(cnt, np.int64), match_idx=g["match_idx"][r, :cnt].astype(np.int64),
                   similarity=g["similarity"][r, :cnt], is_valid=np.ones(cnt, bool))
        got = dict(query_idx=np.zeros(len(ms), np.int64), match_idx=np.array([m.match_idx for m in ms], dtype=np.int64),
                   similarity=np.array([m.similarity for m in ms]), is_valid=np.ones(len(ms), bool))
        parity.compare_candidates(ref, got, k, None)
        assert all(m.query_idx == n for m in ms)
        if with_ts:
            assert not np.any(np.abs(ts[got["match_idx"]] - tq) < gap)
    assert BasePlaceRecognition(descriptor_dim=d, device='cuda').query(desc[0], 0.0) == []     # empty database


@pytest.mark.parametrize("algo", ["lego_loam", "orb_slam3"])
def test_gate_published_counts_gpu(algo):
    """The reference's published gate counts through the gate class
    (results/semantic_gating/{algo}_semantic_analysis.txt:20-22)."""
    from semgate import SemanticLoopClosureGate
    g = np.load(os.path.join(GOLDEN, f"gate_{algo}.npz"))
    i, j = O.spatial_candidates(g["positions"], 2.0, 100)
    total, acc, rej = [int(v) for v in g["published"]]
    gate = SemanticLoopClosureGate(g["floor_labels"], strict_mode=True)
    ok = gate.gate_arrays(i, j)
    st = gate.get_stats()
    assert (st["total_candidates"], st["accepted"], st["rejected_cross_floor"]) == (total, acc, rej)
    want, _ = O.gate_candidates(g["floor_labels"], i, j, True)
    assert np.array_equal(ok, want)
    gate2 = SemanticLoopClosureGate(g["floor_labels"], strict_mode=False)
    gate2.gate_arrays(i, j)
    assert (gate2.stats["accepted"], gate2.stats["rejected_cross_floor"]) == tuple(int(v) for v in g["nonstrict"])
    if algo == "lego_loam":                       # object interface on a slice, order preserved
        cands = [(int(a), int(b), 0.5) for a, b in zip(i[:3000], j[:3000])]
        gate3 = SemanticLoopClosureGate(g["floor_labels"], strict_mode=True)
        valid, rejected = gate3.gate_candidates(cands)
        assert len(valid) == int(want[:3000].sum()) and len(rejected) == 3000 - len(valid)
        assert [c.query_idx for c in valid] == [int(a) for a, w in zip(i[:3000], want[:3000]) if w]
        assert rejected and rejected[0].rejection_reason.startswith("Cross-floor: ")
        one = gate3.gate_candidate(int(i[0]), int(j[0]), 0.9)
        assert one.is_valid == bool(want[0]) and gate3.stats["total_candidates"] == 3001


@pytest.mark.parametrize("M,shift", [(1, 0), (3, 0), (4, 0), (1027, 0), (100003, 0), (4099, 1), (4099, 3)])
def test_gate_candidates_kernel_paths(eng, M, shift):
    """Vector path (4 candidates per thread), its tail, and the scalar path taken for index arrays that
    are only 4-byte aligned; decisions and counters bit-exact against loop_closure_gate.py:89-101."""
    import torch
    rng = np.random.default_rng(M + shift)
    nl = 5000
    fl = rng.choice([1, 2, 4, 5], size=nl).astype(np.int32)
    q = rng.integers(0, nl, size=M + shift).astype(np.int32)
    m = rng.integers(0, nl, size=M + shift).astype(np.int32)
    for mfd in (0, 1):
        v, c = eng.gate_candidates(_t(fl), _t(q)[shift:], _t(m)[shift:], mfd)
        torch.cuda.synchronize()
        want = np.abs(fl[q[shift:]].astype(np.int64) - fl[m[shift:]]) <= mfd
        assert np.array_equal(v.cpu().numpy().astype(bool), want)
        assert c.cpu().numpy().tolist() == [int(want.sum()), int((~want).sum()), 0]
    q[shift] = nl                                                     # out of range -> counted, flagged invalid
    v, c = eng.gate_candidates(_t(fl), _t(q)[shift:], _t(m)[shift:], 0)
    assert int(c[2].item()) == 1 and int(v[0].item()) == 0


@pytest.mark.parametrize("nl,M", [(1, 5000), (2406, 87044), (19163, 600001), (50 * 1024, 600002), (50 * 1024 + 1, 600003), (300, 2999)])
def test_gate_candidates_label_table_in_shared_memory(eng, nl, M):
    """The pair gate keeps the label table in shared memory when it fits beside a block (<= 50 K labels) and there are at
    least 10 candidates per label; otherwise it gathers from global memory.  Both forms, table sizes at the limit,
    negative / too large indices, labels at the int32 extremes: decisions and counters bit-exact (LCG.py:89-101)."""
    import torch
    rng = np.random.default_rng(nl + M)
    fl = rng.choice([1, 2, 4, 5, -2 ** 31, 2 ** 31 - 1], size=nl).astype(np.int32)
    q = rng.integers(0, nl, size=M).astype(np.int32)
    m = rng.integers(0, nl, size=M).astype(np.int32)
    q[::1013] = -1
    m[5::2027] = nl
    m[7::4051] = 2 ** 31 - 1
    bad = (q < 0) | (q >= nl) | (m < 0) | (m >= nl)
    for mfd in (0, 1):
        v, c = eng.gate_candidates(_t(fl), _t(q), _t(m), mfd)
        torch.cuda.synchronize()
        want = (np.abs(fl[np.clip(q, 0, nl - 1)].astype(np.int64) - fl[np.clip(m, 0, nl - 1)]) <= mfd) & ~bad
        assert np.array_equal(v.cpu().numpy().astype(bool), want)
        assert c.cpu().numpy().tolist() == [int(want.sum()), int((~want & ~bad).sum()), int(bad.sum())]


@pytest.mark.parametrize("algo", ["lego_loam", "orb_slam3"])
def test_spatial_candidates_and_gate_end_to_end(algo):
    """Poses -> radius join -> floor gate, all on the GPU, against the reference's published
    counts and the oracle's pair list (orb_slam3_integration.py:167-281)."""
    from semgate import spatial
    g = np.load(os.path.join(GOLDEN, f"gate_{algo}.npz"))
    i, j, d = spatial.detect_loop_closure_candidates_arrays(g["positions"], 2.0, 100)
    oi, oj = O.spatial_candidates(g["positions"], 2.0, 100)
    assert np.array_equal(i, oi) and np.array_equal(j, oj), "candidate pairs differ from the KD-tree join"
    want_d = np.linalg.norm(g["positions"][oi] - g["positions"][oj], axis=1)
    assert np.allclose(d, want_d, rtol=1e-13, atol=0) and d.max() <= 2.0
    a, gate, ok = spatial.apply_floor_gating(i, j, g["floor_labels"], strict_mode=True, max_pairs=5)
    total, acc, rej = [int(v) for v in g["published"]]
    assert (a.total_candidates, a.same_floor_candidates, a.cross_floor_candidates) == (total, acc, rej)
    assert (gate.stats["accepted"], gate.stats["rejected_cross_floor"]) == (acc, rej)
    assert len(a.cross_floor_pairs) == 5 and all(p[2] != p[3] for p in a.cross_floor_pairs)
    a2, gate2, _ = spatial.apply_floor_gating(i, j, g["floor_labels"], strict_mode=False)
    assert (gate2.stats["accepted"], gate2.stats["rejected_cross_floor"]) == tuple(int(v) for v in g["nonstrict"])
    assert a2.same_floor_candidates == acc
    # small cases incl. gap <= 0 and an empty result
    pos = g["positions"][:300]
    for r, gap in ((0.5, 0), (1.0, 1), (3.0, 250), (1e-9, 5)):
        gi, gj, _ = spatial.detect_loop_closure_candidates_arrays(pos, r, gap)
        ri, rj = O.spatial_candidates(pos, r, max(gap, 1))
        assert np.array_equal(gi, ri) and np.array_equal(gj, rj)
    tup = spatial.detect_loop_closure_candidates(pos, 1.0, 50)
    assert all(b - a_ >= 50 and a_ < b for a_, b, _ in tup)


def test_host_abi_find_loop_closures(eng):
    """The reference-facing C entry point on host buffers."""
    from semgate import _native, synthetic
    desc, ts, fl = synthetic.make_case(1200, 256, 3, seed=44)
    p = _native.make_params(k=25, similarity_threshold=0.5, min_time_gap=10.0, max_floor_diff=0)
    q, m, s, v = eng.find_loop_closures_host(desc, ts, fl.astype(np.int32), p)
    got = dict(query_idx=q.astype(np.int64), match_idx=m.astype(np.int64), similarity=s, is_valid=v)
    parity.check_order(got)
    ref = O.find_loop_closures(desc, ts, fl.astype(np.int32), similarity_threshold=0.5, min_time_gap=10.0, k=25)
    parity.compare_candidates(ref, got, 25, 0.5)
    parity.check_decisions_exact(got, ts, fl.astype(np.int32), 10.0, 0)
    # N < 2 -> no candidates (place_recognition.py:864)
    q, m, s, v = eng.find_loop_closures_host(desc[:1], ts[:1], fl[:1].astype(np.int32), p)
    assert len(q) == 0


def test_errors(eng):
    import torch
    from semgate import _native
    with pytest.raises(ValueError):
        _native.make_params(k=0)
    with pytest.raises(ValueError):
        _native.make_params(k=_native.MAX_K_TOTAL + 1)
    x = torch.zeros(4, 64, device="cuda", dtype=torch.bfloat16)
    p = _native.make_params(k=3)
    with pytest.raises(_native.SemgateError):
        eng.gated_topk(x, x, p, q_ts=torch.zeros(4, device="cuda", dtype=torch.float64))   # one-sided timestamps
    from semgate import SemanticPlaceRecognition, SemanticLoopClosureGate
    with pytest.raises(ValueError):
        SemanticPlaceRecognition(vpr_method="nope")
    with pytest.raises(IndexError):
        SemanticLoopClosureGate(np.array([1, 2, 3])).gate_arrays([0], [7])
    assert SemanticPlaceRecognition('mixvpr').find_loop_closures() == []


# --------------------------------------------------------------------------- symmetric all-pairs sweep
def _same_lists(a, b, k, thr):
    """Two sweeps of the same data: identical lists.  (S_ij comes out of a different tile position in the two
    forms; if the tensor core's summation were not symmetric in its operands the scores could differ in the
    last bit, so fall back to the parity rule with a last-bit tolerance before failing.)"""
    if np.array_equal(a["keys"], b["keys"]):
        assert np.array_equal(a["idx"], b["idx"]) and np.array_equal(a["valid"], b["valid"])
        assert np.array_equal(a["count"], b["count"])
        return
    rep = parity.compare_candidates(O.compact(a), O.compact(b), k, thr, tol=2e-6)
    assert rep["max_score_err"] <= 2e-6


SYM_CASES = [
    # n,    d,   k,  thr,   gap,  floors, mode, mfd
    (300,   64,  5,  0.3,   2.0,  3, 0, 0),      # two blocks: diagonal tiles + one off-diagonal tile
    (513,   96,  7,  0.2,   3.0,  4, 0, 0),      # ragged third block, D not a multiple of 64
    (777,   128, 25, 0.5,   10.0, 3, 1, 0),      # mask mode
    (1500,  64,  64, 0.3,   1.0,  3, 0, 1),      # k = 64, non-strict gate
    (2500,  192, 25, 0.45,  10.0, 3, 1, 1),      # mask mode, non-strict
    (5000,  512, 25, 0.5,   10.0, 3, 0, 0),      # C1
    (9001,  64,  10, 0.35,  0.0,  5, 0, -1),     # gap = 0 (self pairs stay), gating off, several super-rows
]


@pytest.mark.parametrize("table", ["1", "0"], ids=["runtable", "superrows"])
@pytest.mark.parametrize("case", SYM_CASES, ids=[f"{c[0]}x{c[1]}k{c[2]}m{c[6]}" for c in SYM_CASES])
def test_symmetric_sweep_equals_full_sweep(eng, case, table, monkeypatch):
    """Queries == database: every similarity computed once and gated in both directions gives the lists of the
    full sweep, and the oracle's.  Both schedules of the triangle: the host-built run table (what sweeps of this
    size use) and the super-row formula (what long sweeps use)."""
    from semgate import synthetic
    monkeypatch.setenv("SEMGATE_SYM_TABLE", table)
    n, d, k, thr, gap, nf, mode, mfd = case
    desc, ts, fl = synthetic.make_case(n, d, nf, seed=n + d)
    fl = fl.astype(np.int32)
    if n == 777:
        fl[::7] = O.FLOOR_NONE                          # unlabelled keyframes pass every gate
    if n == 2500:
        ts = ts[np.random.default_rng(3).permutation(n)].copy()   # unsorted stamps
    kw = dict(k=k, thr=thr, gap=gap, q_ts=ts, db_ts=ts, mfd=mfd, mode=mode, cg=2)
    if mfd >= 0:
        kw.update(q_fl=fl, db_fl=fl)
    full = run_gpu(eng, desc, desc, sym=-1, **kw)
    half = run_gpu(eng, desc, desc, sym=1, **kw)
    assert full["mode"] == 0 and half["mode"] == 1, "the symmetric sweep must have run (and not overflowed)"
    check_padded(half, k)
    _same_lists(full, half, k, thr)
    ref = O.gated_topk(desc, desc, ts, ts, fl if mfd >= 0 else None, fl if mfd >= 0 else None, k=k, threshold=thr,
                       min_time_gap=gap, max_floor_diff=mfd, gate_mode=mode, bf16=True)
    parity.compare_candidates(O.compact(ref), O.compact(half), k, thr, tol=BF16_MODEL_TOL, exact_sets=False)
    parity.check_decisions_exact(O.compact(half), ts, fl if mfd >= 0 else None, gap, mfd)


def test_symmetric_sweep_overflow_falls_back_to_full_sweep(eng):
    """A threshold that admits everything floods the per-keyframe candidate buffers: the overflow flag arms the
    full sweep launched behind the symmetric one, and the lists are still the full sweep's."""
    from semgate import synthetic
    n, d, k = 3000, 64, 25
    desc, ts, fl = synthetic.make_case(n, d, 3, seed=5)
    fl = fl.astype(np.int32)
    kw = dict(k=k, thr=-np.inf, gap=10.0, q_ts=ts, db_ts=ts, q_fl=fl, db_fl=fl, mfd=0, cg=2)
    full = run_gpu(eng, desc, desc, sym=-1, **kw)
    half = run_gpu(eng, desc, desc, sym=1, **kw)
    assert half["mode"] == 2, "the buffers must have overflowed"
    _same_lists(full, half, k, -np.inf)
    # a threshold that lets a few dozen candidates per keyframe through stays inside the buffers
    some = run_gpu(eng, desc, desc, sym=1, **dict(kw, thr=0.3))
    assert some["mode"] == 1
    _same_lists(run_gpu(eng, desc, desc, sym=-1, **dict(kw, thr=0.3)), some, k, 0.3)


@pytest.mark.parametrize("chunks", ["1", "3"])
def test_symmetric_sweep_paced(eng, chunks, monkeypatch):
    """Forced pacing windows: units whose first tiles lie left of the diagonal still arrive on the counters."""
    from semgate import synthetic
    n, d, k = 6000, 256, 25
    desc, ts, fl = synthetic.make_case(n, d, 3, seed=21)
    fl = fl.astype(np.int32)
    kw = dict(k=k, thr=0.5, gap=10.0, q_ts=ts, db_ts=ts, q_fl=fl, db_fl=fl, mfd=0, cg=2)
    full = run_gpu(eng, desc, desc, sym=-1, **kw)
    monkeypatch.setenv("SEMGATE_WINDOW_CHUNKS", chunks)
    monkeypatch.setenv("SEMGATE_SYM_TABLE", "0")          # pacing belongs to the super-row formula
    half = run_gpu(eng, desc, desc, sym=1, **kw)
    assert half["mode"] == 1
    _same_lists(full, half, k, 0.5)


def test_symmetric_sweep_argument_rules(eng):
    import torch
    from semgate import _native
    x = eng.normalize_cast(torch.randn(600, 64, device="cuda"))
    y = x.clone()
    ts = torch.arange(600, device="cuda", dtype=torch.float64)
    with pytest.raises(_native.SemgateError):           # different matrices
        eng.gated_topk(x, y, _native.make_params(k=5, cta_group=2, symmetric=1), q_ts=ts, db_ts=ts)
    with pytest.raises(_native.SemgateError):           # single-CTA tiles
        eng.gated_topk(x, x, _native.make_params(k=5, cta_group=1, symmetric=1), q_ts=ts, db_ts=ts)
    with pytest.raises(_native.SemgateError):           # different stamps
        eng.gated_topk(x, x, _native.make_params(k=5, cta_group=2, symmetric=1), q_ts=ts, db_ts=ts.clone())
    with pytest.raises(_native.SemgateError):           # parts only split a symmetric sweep
        eng.gated_topk(x, x, _native.make_params(k=5, cta_group=2, part_index=0, part_count=2), q_ts=ts, db_ts=ts)
    with pytest.raises(_native.SemgateError):
        eng.gated_topk(x, x, _native.make_params(k=5, cta_group=2, symmetric=1, part_index=2, part_count=2), q_ts=ts, db_ts=ts)
    auto = dict(k=5, cta_group=2, similarity_threshold=0.3)
    eng.gated_topk(x, x, _native.make_params(**auto), q_ts=ts, db_ts=ts)          # auto: too small to pay -> full sweep
    assert eng.last_sweep_mode() == (0, 9)
    eng.set_option("symmetric", 1)                                                # whenever the arguments allow
    try:
        eng.gated_topk(x, y, _native.make_params(**auto), q_ts=ts, db_ts=ts)      # different matrices: full sweep
        assert eng.last_sweep_mode()[0] == 0
        eng.gated_topk(x, x, _native.make_params(**auto), q_ts=ts, db_ts=ts)
        assert eng.last_sweep_mode() == (1, 6)                                    # 3 blocks: 6 tiles of the 9
        eng.set_option("symmetric", -1)
        eng.gated_topk(x, x, _native.make_params(**auto), q_ts=ts, db_ts=ts)
        assert eng.last_sweep_mode() == (0, 9)
    finally:
        eng.set_option("symmetric", 0)
    # auto by size: a long, tensor-bound all-pairs sweep takes the symmetric schedule by itself
    big = eng.normalize_cast(torch.randn(8200, 1024, device="cuda"))
    eng.gated_topk(big, big, _native.make_params(k=5, similarity_threshold=0.3))
    assert eng.last_sweep_mode()[0] == 1
    eng.gated_topk(big[:8000], big[:8000], _native.make_params(k=5, similarity_threshold=0.3))
    assert eng.last_sweep_mode()[0] == 0


@pytest.mark.parametrize("table", ["1", "0"], ids=["runtable", "superrows"])
@pytest.mark.parametrize("G", [2, 3, 8])
def test_symmetric_sweep_parts_merge_to_full_sweep(eng, G, table, monkeypatch):
    """The multi-GPU form on one GPU: the G parts of the tile triangle, swept one after the other and merged by
    K3, give the lists of the full sweep."""
    import torch
    from semgate import _native, synthetic
    monkeypatch.setenv("SEMGATE_SYM_TABLE", table)
    n, d, k = 12000, 128, 25
    desc, ts, fl = synthetic.make_case(n, d, 4, seed=77)
    xb = eng.normalize_cast(_t(desc, torch.float32))
    tts, tfl = _t(ts, torch.float64), _t(fl.astype(np.int32), torch.int32)
    common = dict(k=k, similarity_threshold=0.5, min_time_gap=10.0, max_floor_diff=0, cta_group=2)
    full = eng.gated_topk(xb, xb, _native.make_params(symmetric=-1, **common), q_ts=tts, db_ts=tts, q_floor=tfl, db_floor=tfl,
                          want_keys=True)
    keys, tiles = [], 0
    for g in range(G):
        p = _native.make_params(symmetric=1, part_index=g, part_count=G, **common)
        r = eng.gated_topk(xb, xb, p, q_ts=tts, db_ts=tts, q_floor=tfl, db_floor=tfl, want_lists=False, want_keys=True)
        mode, t = eng.last_sweep_mode()
        assert mode == 1
        tiles += t
        keys.append(r.keys.clone())
    nb = (n + 255) // 256
    assert tiles == nb * (nb + 1) // 2, "the parts must tile the triangle exactly once"
    merged = eng.merge_topk(torch.stack(keys), k, q_floor=tfl, db_floor_all=tfl, max_floor_diff=0, want_keys=True)
    torch.cuda.synchronize()
    a = dict(keys=full.keys.cpu().numpy(), idx=full.idx.cpu().numpy().astype(np.int64), valid=full.valid.cpu().numpy().astype(bool),
             count=full.count.cpu().numpy(), scores=full.scores.cpu().numpy())
    b = dict(keys=merged.keys.cpu().numpy(), idx=merged.idx.cpu().numpy().astype(np.int64),
             valid=merged.valid.cpu().numpy().astype(bool), count=merged.count.cpu().numpy(), scores=merged.scores.cpu().numpy())
    _same_lists(a, b, k, 0.5)


# --------------------------------------------------------------------------- torch.ops.semgate
def test_torch_ops_match_the_engine(eng):
    """The PyTorch operators end in the same kernels as the ctypes engine: identical outputs, including the
    symmetric sweep taken automatically for aliased arguments, a sharded sweep merged by merge_topk, and compact."""
    import torch
    from semgate import _native, ops, synthetic
    sg = ops.load()
    n, d, k = 4500, 200, 25
    desc, ts, fl = synthetic.make_case(n, d, 3, seed=9)
    x = _t(desc, torch.float32)
    tts, tfl = _t(ts, torch.float64), _t(fl.astype(np.int32), torch.int32)
    xb = sg.normalize_cast(x)
    assert torch.equal(xb, eng.normalize_cast(x)) and xb.shape == (n, 256)
    for half in (torch.float16, torch.bfloat16):      # an extractor under autocast hands over half-precision CUDA tensors
        assert torch.equal(sg.normalize_cast(x.to(half)), eng.normalize_cast(x.to(half).float()))
    with pytest.raises(RuntimeError):
        sg.normalize_cast(x.double())
    p = _native.make_params(k=k, similarity_threshold=0.5, min_time_gap=10.0, max_floor_diff=0)
    want_ = eng.gated_topk(xb, xb, p, q_ts=tts, db_ts=tts, q_floor=tfl, db_floor=tfl, want_keys=True)
    sc, ix, va, ct, keys = sg.gated_topk(xb, xb, tfl, tfl, tts, tts, 10.0, 0.5, k, 0, 0, 0)
    torch.cuda.synchronize()
    for a, b in ((sc, want_.scores), (ix, want_.idx), (va, want_.valid), (ct, want_.count), (keys, want_.keys)):
        assert torch.equal(a, b)
    # two database shards + merge == whole
    parts = []
    for lo, hi in ((0, 2000), (2000, n)):
        parts.append(sg.gated_topk(xb, xb[lo:hi], tfl, tfl[lo:hi].contiguous(), tts, tts[lo:hi].contiguous(), 10.0, 0.5, k, 0, 0, lo)[4])
    msc, mix, mva, mct, _ = sg.merge_topk(torch.stack(parts), tfl, tfl, 0)
    assert torch.equal(mix, ix) and torch.equal(msc, sc) and torch.equal(mva, va) and torch.equal(mct, ct)
    oq, om, os_, ov, total = sg.compact(sc, ix, va, ct)
    eq, em, es, ev, et = eng.compact(want_)
    t = int(total.item())
    assert t == int(et.item()) == int(ct.sum().item())
    assert torch.equal(oq[:t], eq[:t]) and torch.equal(om[:t], em[:t]) and torch.equal(os_[:t], es[:t]) and torch.equal(ov[:t], ev[:t])
    # no timestamps / labels, query form (threshold off)
    q1 = sg.gated_topk(xb[:3].contiguous(), xb, None, None, None, None, 10.0, float("-inf"), 5, -1, 0, 0)
    r1 = eng.gated_topk(xb[:3].contiguous(), xb, _native.make_params(k=5))
    assert torch.equal(q1[1], r1.idx) and torch.equal(q1[0], r1.scores)
    with pytest.raises(RuntimeError):
        sg.gated_topk(xb.float(), xb, None, None, None, None, 10.0, 0.5, k, 0, 0, 0)      # wrong dtype


# --------------------------------------------------------------------------- full-size properties (BASELINE config 2)
@pytest.mark.parametrize("cg", [1, 2, 4])
def test_full_size_properties_c2(eng, cg):
    """20k x 4096-d all-pairs sweep: sampled rows against the oracle + structural properties."""
    import torch
    from semgate import _native, synthetic
    n, d, k = 20000, 4096, 25
    x = synthetic.make_descriptors_device(n, d, "cuda", seed=0)
    ts = synthetic.make_timestamps(n)
    fl = synthetic.make_floors(n, 3).astype(np.int32)
    xb = eng.normalize_cast(x)
    p = _native.make_params(k=k, similarity_threshold=0.5, min_time_gap=10.0, max_floor_diff=0, cta_group=cg)
    tts, tfl = _t(ts), _t(fl)
    r = eng.gated_topk(xb, xb, p, q_ts=tts, db_ts=tts, q_floor=tfl, db_floor=tfl)
    r2 = eng.gated_topk(xb, xb, p, q_ts=tts, db_ts=tts, q_floor=tfl, db_floor=tfl)
    torch.cuda.synchronize()
    got = dict(scores=r.scores.cpu().numpy(), idx=r.idx.cpu().numpy().astype(np.int64),
               valid=r.valid.cpu().numpy().astype(bool), count=r.count.cpu().numpy())
    check_padded(got, k)
    assert torch.equal(r.idx, r2.idx) and torch.equal(r.scores, r2.scores), "sweep must be deterministic"
    c = O.compact(got)
    parity.check_decisions_exact(c, ts, fl, 10.0, 0)
    assert c["similarity"].min() >= np.float32(0.5)
    # symmetry of the all-pairs sweep: (i,j) above threshold and not cut by k appears as (j,i)
    rows = np.random.default_rng(0).choice(n, 96, replace=False)
    xs = x[torch.from_numpy(rows).cuda()].cpu().numpy()
    xa = x.cpu().numpy()
    ref = O.gated_topk(xs, xa, ts[rows], ts, fl[rows], fl, k=k, threshold=0.5, min_time_gap=10.0, max_floor_diff=0)
    sub = dict(scores=got["scores"][rows], idx=got["idx"][rows], valid=got["valid"][rows], count=got["count"][rows])
    rep = parity.compare_candidates(O.compact(ref), O.compact(sub), k, 0.5)
    assert rep["max_score_err"] <= parity.SCORE_TOL
    oq, om, os_, ov, total = eng.compact(r)
    t = int(total.item())
    assert t == int(got["count"].sum())
    flat = dict(query_idx=oq[:t].cpu().numpy(), match_idx=om[:t].cpu().numpy(), similarity=os_[:t].cpu().numpy(),
                is_valid=ov[:t].cpu().numpy().astype(bool))
    parity.check_order(flat)
    assert np.array_equal(flat["match_idx"], c["match_idx"]) and np.array_equal(flat["is_valid"], c["is_valid"])


def _sampled_rows_reference(xq_bf16, xdb_bf16, rows, ts_q, ts_db, fl_q, fl_db, k, thr, gap):
    """fp32 torch reference of the same op on a sample of query rows (the full Q x N matrix of the
    large configs cannot exist): similarities of the bf16 operands accumulated in fp32, window in fp64,
    top-k, threshold, floor flag — the order of place_recognition.py:882-899."""
    import torch
    out = {"query_idx": [], "match_idx": [], "similarity": [], "is_valid": []}
    r = torch.from_numpy(rows).cuda()
    q = xq_bf16[r].float()
    sims = torch.empty((len(rows), xdb_bf16.shape[0]), dtype=torch.float32, device="cuda")
    step = 65536
    for s0 in range(0, xdb_bf16.shape[0], step):
        sims[:, s0:s0 + step] = q @ xdb_bf16[s0:s0 + step].float().T
    tq = torch.from_numpy(ts_q[rows]).cuda()
    tdb = torch.from_numpy(ts_db).cuda()
    sims[(tdb[None, :] - tq[:, None]).abs() < gap] = -float("inf")
    top_s, top_i = torch.topk(sims, k, dim=1)
    top_s, top_i = top_s.cpu().numpy(), top_i.cpu().numpy()
    for a, row in enumerate(rows):
        for s, j in zip(top_s[a], top_i[a]):
            if np.isfinite(s) and s >= np.float32(thr):
                out["query_idx"].append(int(row)); out["match_idx"].append(int(j)); out["similarity"].append(float(s))
                out["is_valid"].append(bool(O.floor_ok(fl_q[row], fl_db[j], 0)))
    return {key: np.asarray(v) for key, v in out.items()}


@pytest.mark.parametrize("cfg", ["c3", "c5"])
def test_full_size_properties_large_configs(eng, cfg):
    """BASELINE configs 3 (10k x 100k x 8448-d) and 5 (1M x 1M x 4096-d, 16 floors) at full size: structural
    invariants of every list, bit-exact decisions on every returned pair, determinism, and 64 sampled query
    rows against an fp32 torch reference of the same op (scores within the bf16-model tolerance, sets equal up
    to the boundary rule)."""
    import torch
    from semgate import _native, synthetic
    Q, N, D, F = (10_000, 100_000, 8448, 4) if cfg == "c3" else (1_000_000, 1_000_000, 4096, 16)
    k, thr, gap = 25, 0.5, 10.0
    dp = _native.pad_dim(D)
    xb = torch.empty((N, dp), dtype=torch.bfloat16, device="cuda")
    g = torch.Generator(device="cuda"); g.manual_seed(7)
    places = max(8, N // 20)
    anchors = torch.randn((places, D), generator=g, device="cuda")
    step = max(1024, (1 << 27) // D)
    for s0 in range(0, N, step):
        e0 = min(N, s0 + step)
        pid = torch.randint(0, places, (e0 - s0,), generator=g, device="cuda")
        eng.normalize_cast(anchors[pid] + 0.6 * torch.randn((e0 - s0, D), generator=g, device="cuda"), out=xb[s0:e0])
    del anchors
    ts = synthetic.make_timestamps(N)
    fl = synthetic.make_floors(N, F).astype(np.int32)
    tts, tfl = _t(ts), _t(fl)
    p = _native.make_params(k=k, similarity_threshold=thr, min_time_gap=gap, max_floor_diff=0)
    qb, qts, qfl = xb[:Q], tts[:Q].contiguous(), tfl[:Q].contiguous()
    r = eng.gated_topk(qb, xb, p, q_ts=qts, db_ts=tts, q_floor=qfl, db_floor=tfl, want_keys=True)
    torch.cuda.synchronize()
    if cfg == "c3":
        r2 = eng.gated_topk(qb, xb, p, q_ts=qts, db_ts=tts, q_floor=qfl, db_floor=tfl, want_keys=True)
        assert torch.equal(r.keys, r2.keys), "sweep must be deterministic"
    else:
        # the aliased arguments above take the symmetric sweep; the full-matrix kernel at 1M x 1M must give the
        # same lists bit for bit (the full-matrix form was untested at this size)
        assert eng.last_sweep_mode()[0] == 1
        pf = _native.make_params(k=k, similarity_threshold=thr, min_time_gap=gap, max_floor_diff=0, symmetric=-1)
        r2 = eng.gated_topk(qb, xb, pf, q_ts=qts, db_ts=tts, q_floor=qfl, db_floor=tfl, want_keys=True)
        torch.cuda.synchronize()
        assert eng.last_sweep_mode()[0] == 0
        assert torch.equal(r.keys, r2.keys), "symmetric and full-matrix sweeps differ at 1M x 1M"
        del r2
    # structure, on the device (1M x 25 lists)
    pos = torch.arange(k, device="cuda")[None, :]
    filled = pos < r.count[:, None]
    assert bool(((r.idx >= 0) == filled).all()) and bool((r.scores[filled] >= thr).all())
    assert bool((r.scores[:, 1:][filled[:, 1:]] <= r.scores[:, :-1][filled[:, 1:]]).all()), "scores not descending"
    assert bool((r.idx[filled] < N).all()) and bool((r.valid[~filled] == 0).all())
    # decisions, bit-exact, on every returned pair
    qi = torch.arange(Q, device="cuda")[:, None].expand(Q, k)[filled]
    mi = r.idx[filled].long()
    assert not bool(((tts[mi] - tts[qi]).abs() < gap).any()), "a returned pair lies inside the exclusion window"
    assert bool(((tfl[qi] == tfl[mi]) == (r.valid[filled] != 0)).all()), "floor-gate bits differ"
    total = int(r.count.sum().item())
    assert total > Q                                   # the synthetic places give every keyframe revisits
    # sampled rows against the fp32 torch reference
    rows = np.sort(np.random.default_rng(3).choice(Q, 64, replace=False))
    ref = _sampled_rows_reference(qb, xb, rows, ts[:Q], ts, fl[:Q], fl, k, thr, gap)
    rr = torch.from_numpy(rows).cuda()
    sub = dict(scores=r.scores[rr].cpu().numpy(), idx=r.idx[rr].cpu().numpy().astype(np.int64),
               valid=r.valid[rr].cpu().numpy().astype(bool), count=r.count[rr].cpu().numpy())
    got = O.compact(sub)
    got["query_idx"] = rows[got["query_idx"]]
    rep = parity.compare_candidates(ref, got, k, thr, tol=BF16_MODEL_TOL)
    assert rep["max_score_err"] <= BF16_MODEL_TOL


# --------------------------------------------------------------------------- K5: CricaVPR cross-correlation re-rank
RER = sorted(glob.glob(os.path.join(GOLDEN, "rerank_*.npz")))


@pytest.mark.parametrize("path", RER, ids=[os.path.basename(p)[:-4] for p in RER])
def test_rerank_golden(path):
    """compute_cross_correlation_score / rerank_candidates (place_recognition.py:669-757) against the
    reference's own outputs: scores within 2e-3, re-ranked lists equal up to that tolerance."""
    from semgate import CricaVPR, synthetic
    g = np.load(path)
    n, patches, dim, seed, top_k, ncand, missing = [int(v) for v in g["params"]]
    feats, _ = synthetic.make_local_features(n, patches, dim, seed)
    vpr = CricaVPR(device='cuda', use_reranking=True)
    for i in range(n):
        vpr.add_image(np.zeros(vpr.descriptor_dim, np.float32), float(i), 1,
                      local_features=None if i == missing else feats[i])
    assert (missing in vpr._feature_cache) is False and len(vpr._feature_cache) == n - (1 if missing >= 0 else 0)
    for r, q in enumerate(g["query_idx"].tolist()):
        cands = [(int(c), float(s)) for c, s in zip(g["cand_idx"][r], g["cand_sim"][r])]
        m = np.array([c for c, _ in cands])
        cross, comb = vpr._pair_scores(np.full(len(m), q), m, g["cand_sim"][r])
        ref_cross = g["cross"][r]
        has = m != missing
        assert np.all(np.isnan(cross[~has])) and np.array_equal(comb[~has], g["cand_sim"][r][~has])
        assert np.max(np.abs(cross[has] - ref_cross[has])) <= parity.SCORE_TOL
        want16 = np.array([O.cross_correlation_score(feats[q], feats[c], bf16=True) for c in m[has]])
        assert np.max(np.abs(cross[has] - want16)) <= BF16_MODEL_TOL
        rr = vpr.rerank_candidates(q, cands, top_k=top_k)
        cnt = int(g["out_count"][r])
        assert len(rr) == cnt
        ref_list = list(zip(g["out_idx"][r, :cnt].tolist(), g["out_score"][r, :cnt].tolist()))
        ref_all = dict(O.rerank_candidates({i: feats[i] for i in range(n) if i != missing}, q, cands, top_k=len(cands)))
        for (gi, gs), (ri, rs) in zip(rr, ref_list):
            assert abs(gs - ref_all[gi]) <= parity.SCORE_TOL          # every returned score is right
            assert gi == ri or abs(ref_all[gi] - rs) <= 2 * parity.SCORE_TOL   # order differs only inside the tolerance
    # single-pair entry point and the untouched-order cases
    s = vpr.compute_cross_correlation_score(feats[0], feats[1][None])
    assert abs(s - float(O.cross_correlation_score(feats[0], feats[1]))) <= parity.SCORE_TOL
    assert vpr.rerank_candidates(10 ** 6, [(1, 0.5), (2, 0.9)], top_k=1) == [(1, 0.5)]
    vpr.use_reranking = False
    assert vpr.rerank_candidates(0, [(1, 0.5), (2, 0.9)], top_k=5) == [(1, 0.5), (2, 0.9)]


def test_rerank_cluster_lockstep_cases(eng, monkeypatch):
    """K5 runs two consecutive pairs per 2-CTA cluster in lock-step: same query (multicast operand),
    different queries (private operands), one or both pairs without cached features (the idle CTA only
    keeps the stage ring turning), an odd tail; results equal the single-CTA form bit for bit and the oracle."""
    import torch
    from semgate import synthetic
    n, P, D = 12, 200, 128
    feats, _ = synthetic.make_local_features(n, P, D, seed=4)
    fb = eng.normalize_cast(_t(feats.reshape(n * P, D))).view(n, P, -1)
    pairs = [(0, 1), (0, 2),        # shared query
             (1, 3), (2, 3),        # different queries
             (5, -1), (5, 6),       # first pair of the group invalid
             (4, 7), (99, 7),       # second invalid (index beyond the store)
             (-1, 2), (-1, 3),      # both invalid: the group is skipped
             (8, 9), (8, 10), (8, 11),   # odd tail: last group has one pair
             ]
    q = np.array([a for a, _ in pairs], np.int32)
    m = np.array([b for _, b in pairs], np.int32)
    g = np.linspace(0.1, 0.9, len(pairs)).astype(np.float32)
    outs = {}
    for cl in ("2", "1"):
        monkeypatch.setenv("SEMGATE_RERANK_CLUSTER", cl)
        cross, comb = eng.rerank_scores(fb, _t(q), _t(m), _t(g))
        torch.cuda.synchronize()
        outs[cl] = (cross.cpu().numpy(), comb.cpu().numpy())
    monkeypatch.delenv("SEMGATE_RERANK_CLUSTER")
    assert np.array_equal(outs["2"][0], outs["1"][0], equal_nan=True) and np.array_equal(outs["2"][1], outs["1"][1])
    cross, comb = outs["2"]
    for i, (a, b) in enumerate(pairs):
        if 0 <= a < n and 0 <= b < n:
            want = float(O.cross_correlation_score(feats[a], feats[b], bf16=True))
            assert abs(cross[i] - want) <= BF16_MODEL_TOL, (i, a, b, cross[i], want)
            assert abs(comb[i] - (0.5 * g[i] + 0.5 * want)) <= BF16_MODEL_TOL
        else:
            assert np.isnan(cross[i]) and comb[i] == g[i]


@pytest.mark.parametrize("P,D", [(129, 64), (200, 128), (257, 64), (280, 192), (288, 64), (300, 128), (529, 768), (544, 64), (545, 128),
                                 (800, 128), (1030, 64)])
def test_rerank_pair_form(eng, monkeypatch, P, D):
    """K5 pair form (one CTA pair per candidate pair, cta_group::2 tiles; left-over candidate patches as strip MMAs when
    P = 256 a + r, r <= 32) against the single-CTA form and an fp32 reference of the same op on the same bf16 rows, over
    patch counts that hit every tiling case: one / several pair tiles, with and without a strip, ragged last n-tile;
    pairs without cached features; a second launch over the same buffers (the merge arrays come back clean)."""
    import torch
    n, M = 20, 203
    g = torch.Generator(device="cuda").manual_seed(P * 7 + D)
    x = torch.randn((n * P, D), device="cuda", generator=g) + 0.3
    fb = eng.normalize_cast(x).view(n, P, -1)
    qi = torch.randint(0, n, (M,), device="cuda", dtype=torch.int32, generator=g)
    mi = torch.randint(0, n, (M,), device="cuda", dtype=torch.int32, generator=g)
    qi[7] = -1
    mi[11] = n + 3
    qi[M - 1] = -1
    qi[20:45] = 3
    gs = torch.rand((M,), device="cuda", generator=g)
    monkeypatch.delenv("SEMGATE_RERANK_CLUSTER", raising=False)
    cross, comb = eng.rerank_scores(fb, qi, mi, gs)
    cross2, _ = eng.rerank_scores(fb, qi, mi, gs)
    monkeypatch.setenv("SEMGATE_RERANK_CLUSTER", "1")
    cross1, comb1 = eng.rerank_scores(fb, qi, mi, gs)
    torch.cuda.synchronize()
    monkeypatch.delenv("SEMGATE_RERANK_CLUSTER")
    ff = fb.float()
    valid = ((qi >= 0) & (mi >= 0) & (qi < n) & (mi < n)).cpu().numpy()
    ref = np.full(M, np.nan, np.float32)
    for i in np.nonzero(valid)[0]:
        c = ff[int(qi[i])] @ ff[int(mi[i])].T
        ref[i] = float(torch.sqrt(c.max(dim=1).values.mean() * c.max(dim=0).values.mean()))
    cross, comb, cross1, cross2 = cross.cpu().numpy(), comb.cpu().numpy(), cross1.cpu().numpy(), cross2.cpu().numpy()
    assert np.array_equal(np.isnan(cross), ~valid) and np.array_equal(np.isnan(cross1), ~valid)
    assert np.array_equal(cross, cross2, equal_nan=True)
    assert np.max(np.abs(cross[valid] - ref[valid])) <= 2e-6          # the maxima are exact; only the order of the fp32 sums differs
    assert np.max(np.abs(cross[valid] - cross1[valid])) <= 2e-6
    gsn = gs.cpu().numpy()
    assert np.array_equal(comb[~valid], gsn[~valid]) and np.allclose(comb[valid], 0.5 * gsn[valid] + 0.5 * cross[valid], atol=1e-6)


def test_rerank_batch_dinov2_shape(eng):
    """529 patches x 768-d (DINOv2 at 322x322): a batch of pairs against the oracle, and the batched
    per-query selection against per-query Python sorting."""
    import torch
    from semgate import CricaVPR, synthetic
    n, P, D, kc, top_k = 40, 529, 768, 12, 5
    feats, _ = synthetic.make_local_features(n, P, D, seed=9)
    vpr = CricaVPR(device='cuda')
    for i in range(n):
        vpr.add_image(np.zeros(vpr.descriptor_dim, np.float32), float(i), 1, local_features=feats[i] if i % 9 != 4 else None)
    rng = np.random.default_rng(0)
    Q = 16
    query_idx = rng.choice(n, Q, replace=False)
    cand = np.stack([rng.choice(n, kc, replace=False) for _ in range(Q)]).astype(np.int32)
    gsim = rng.uniform(0.3, 0.9, size=(Q, kc)).astype(np.float32)
    count = rng.integers(3, kc + 1, size=Q).astype(np.int32)
    oi, os_, oc = vpr.rerank_batch(query_idx, cand, gsim, count, top_k=top_k)
    cache = {i: feats[i] for i in vpr._feature_cache}
    for r in range(Q):
        cl = [(int(c), float(s)) for c, s in zip(cand[r, :count[r]], gsim[r, :count[r]])]
        ref = O.rerank_candidates(cache, int(query_idx[r]), cl, top_k=top_k, bf16=True)
        full = dict(O.rerank_candidates(cache, int(query_idx[r]), cl, top_k=len(cl), bf16=True))
        assert oc[r] == len(ref)
        for t in range(oc[r]):
            assert abs(os_[r, t] - full[int(oi[r, t])]) <= BF16_MODEL_TOL
            assert int(oi[r, t]) == ref[t][0] or abs(full[int(oi[r, t])] - ref[t][1]) <= 2 * BF16_MODEL_TOL
        assert np.all(oi[r, oc[r]:] == -1)


# --------------------------------------------------------------------------- round 2: k > 64, row-slice merge, one-call sweep, resync
@pytest.mark.parametrize("k,cg,Q,N,D", [(100, 1, 300, 900, 128), (100, 2, 300, 900, 128), (65, 1, 70, 700, 64),
                                        (130, 2, 513, 1400, 192), (200, 0, 3, 2500, 256), (1000, 1, 40, 1100, 64)])
def test_k_above_64_runs_as_several_sweeps(eng, k, cg, Q, N, D):
    """The reference's k is any integer (place_recognition.py:853,888); above 64 the library runs ceil(k/64) sweeps,
    each admitting only keys below the last key of the sweep before.  Against the oracle at the same k: lists,
    scores, flags; rows with fewer than k admissible candidates end early and stay consistent."""
    from semgate import synthetic
    desc, ts, fl = synthetic.make_case(max(Q, N), D, 3, seed=k + Q)
    fl32 = fl.astype(np.int32)
    thr = -0.05 if k < 1000 else -np.inf
    got = run_gpu(eng, desc[:Q], desc[:N], k, thr, 3.0, ts[:Q], ts[:N], fl32[:Q], fl32[:N], mfd=0, cg=cg)
    check_padded(got, k)
    assert got["count"].max() > 64, "the case must exercise more than one pass"
    ref16 = O.gated_topk(desc[:Q], desc[:N], ts[:Q], ts[:N], fl32[:Q], fl32[:N], k=k, threshold=thr, min_time_gap=3.0,
                         max_floor_diff=0, bf16=True)
    rep = parity.compare_candidates(O.compact(ref16), O.compact(got), k, thr, tol=BF16_MODEL_TOL)
    assert rep["max_score_err"] < BF16_MODEL_TOL
    ref32 = O.gated_topk(desc[:Q], desc[:N], ts[:Q], ts[:N], fl32[:Q], fl32[:N], k=k, threshold=thr, min_time_gap=3.0,
                         max_floor_diff=0)
    parity.compare_candidates(O.compact(ref32), O.compact(got), k, thr, tol=parity.SCORE_TOL)
    parity.check_decisions_exact(O.compact(got), ts[:N], fl32[:N], 3.0, 0, q_ts=ts[:Q], q_floors=fl32[:Q])
    assert np.array_equal(got["keys"], O.pack_keys(got["scores"], got["idx"]))
    # the first 64 columns are the single-sweep answer, bit for bit
    # (a handful of query rows would take the streaming kernel, whose fp32 sums round differently: pin the tile form)
    one = run_gpu(eng, desc[:Q], desc[:N], 64, thr, 3.0, ts[:Q], ts[:N], fl32[:Q], fl32[:N], mfd=0, cg=cg or 1)
    assert np.array_equal(one["idx"], got["idx"][:, :64]) and np.array_equal(one["scores"], got["scores"][:, :64])
    if k == 100:
        # a threshold that leaves some rows short of 64, some between 64 and k, some full: later passes must add
        # nothing to a row that already ran out of candidates
        thr2 = 0.12
        got2 = run_gpu(eng, desc[:Q], desc[:N], k, thr2, 3.0, ts[:Q], ts[:N], fl32[:Q], fl32[:N], mfd=0, cg=cg)
        check_padded(got2, k)
        c = got2["count"]
        assert c.min() < 64 < c.max()
        ref2 = O.gated_topk(desc[:Q], desc[:N], ts[:Q], ts[:N], fl32[:Q], fl32[:N], k=k, threshold=thr2, min_time_gap=3.0,
                            max_floor_diff=0, bf16=True)
        parity.compare_candidates(O.compact(ref2), O.compact(got2), k, thr2, tol=BF16_MODEL_TOL)


def test_k_above_64_through_the_mirrored_classes_and_host_abi(eng):
    from semgate import SemanticPlaceRecognition, PlaceDescriptor, _native, synthetic
    n, d, k = 1200, 128, 100
    desc, ts, fl = synthetic.make_case(n, d, 3, seed=5)
    fl32 = fl.astype(np.int32)
    spr = SemanticPlaceRecognition('mixvpr', 'cuda', similarity_threshold=0.05, min_time_gap=4.0, descriptor_dim=d)
    for i in range(n):
        spr.vpr.descriptors.append(PlaceDescriptor(float(ts[i]), desc[i], floor_label=int(fl[i])))
    arr = spr.find_loop_closures_arrays(enable_floor_gating=True, k=k)
    got = dict(query_idx=arr.query_idx.astype(np.int64), match_idx=arr.match_idx.astype(np.int64), similarity=arr.similarity,
               is_valid=arr.is_valid)
    ref = O.find_loop_closures(desc, ts, fl32, similarity_threshold=0.05, min_time_gap=4.0, k=k)
    parity.compare_candidates(ref, got, k, 0.05)
    parity.check_order(got)
    assert np.bincount(got["query_idx"]).max() > 64
    # host-buffer C ABI, same k
    p = _native.make_params(k=k, similarity_threshold=0.05, min_time_gap=4.0, max_floor_diff=0)
    q, m, s, v = eng.find_loop_closures_host(desc, ts, fl32, p)
    assert np.array_equal(q, arr.query_idx) and np.array_equal(m, arr.match_idx) and np.array_equal(s, arr.similarity)
    assert np.array_equal(v, arr.is_valid)
    # query(): k above the database size comes back with every admissible keyframe, like the reference's slice
    res = spr.vpr.query(desc[7], timestamp=float(ts[7]), k=1000, min_time_gap=4.0)
    want = int((np.abs(ts - ts[7]) >= 4.0).sum())
    assert len(res) == min(want, 1000)
    sims = np.array([r.similarity for r in res])
    assert np.all(np.diff(sims) <= 0)
    small = SemanticPlaceRecognition('mixvpr', 'cuda', similarity_threshold=-1.0, min_time_gap=0.0, descriptor_dim=d)
    for i in range(90):
        small.vpr.descriptors.append(PlaceDescriptor(float(ts[i]), desc[i], floor_label=int(fl[i])))
    every = small.vpr.query(desc[200], timestamp=None, k=5000)          # k beyond the database: all 90 come back
    assert len(every) == 90 and len({m.match_idx for m in every}) == 90
    assert spr.find_loop_closures(k=0) == []
    with pytest.raises(ValueError):
        spr.find_loop_closures(k=-1)


def test_accumulate_cannot_flag_other_slices(eng):
    """the seeded lists of an accumulating sweep hold indices of other database slices; their floor flags
    cannot come from this slice's labels -> SEMGATE_EINVAL instead of an out-of-bounds read."""
    import torch
    from semgate import _native, synthetic
    desc, ts, fl = synthetic.make_case(1500, 64, 3, seed=2)
    xb = eng.normalize_cast(_t(desc))
    tts, tfl = _t(ts), _t(fl.astype(np.int32))
    kw = dict(k=10, similarity_threshold=0.2, min_time_gap=5.0, max_floor_diff=0)
    first = eng.gated_topk(xb, xb[:800], _native.make_params(**kw), q_ts=tts, db_ts=tts[:800].contiguous(), q_floor=tfl,
                           db_floor=tfl[:800].contiguous(), want_keys=True)
    with pytest.raises(_native.SemgateError) as e:
        eng.gated_topk(xb, xb[800:], _native.make_params(db_index_offset=800, accumulate=True, **kw), q_ts=tts,
                       db_ts=tts[800:].contiguous(), q_floor=tfl, db_floor=tfl[800:].contiguous(), want_lists=True, keys=first.keys)
    assert e.value.code == _native.EINVAL
    # without gating there is nothing to flag: allowed, and equal to the whole sweep
    kw2 = dict(kw, max_floor_diff=-1)
    a = eng.gated_topk(xb, xb[:800], _native.make_params(**kw2), q_ts=tts, db_ts=tts[:800].contiguous(), want_keys=True)
    b = eng.gated_topk(xb, xb[800:], _native.make_params(db_index_offset=800, accumulate=True, **kw2), q_ts=tts,
                       db_ts=tts[800:].contiguous(), want_lists=True, keys=a.keys)
    whole = eng.gated_topk(xb, xb, _native.make_params(**kw2), q_ts=tts, db_ts=tts)
    torch.cuda.synchronize()
    assert torch.equal(b.idx, whole.idx) and torch.equal(b.scores, whole.scores)


def test_merge_topk_peers_rows_and_flags(eng):
    """The row-slice form of the peer merge (every rank merges only its own rows and ORs the ranks' overflow
    flags): equals the same rows of the gathered merge; here all "peers" live on this GPU."""
    import torch
    rng = np.random.default_rng(9)
    for G, Q, k, flags in ((2, 1001, 25, (0, 0)), (8, 333, 25, (0, 0, 0, 1, 0, 0, 0, 0)), (4, 257, 64, (0, 2, 0, 0)), (3, 50, 7, (0, 0, 5))):
        sc = rng.uniform(-1, 1, size=(G, Q, k)).astype(np.float32)
        ix = (rng.permutation(G * Q * k).reshape(G, Q, k) % (2 ** 20)).astype(np.int64)
        ix[rng.uniform(size=ix.shape) > 0.8] = -1
        keys = torch.from_numpy(O.pack_keys(sc, ix)).cuda()
        parts = []
        for g in range(G):
            buf = torch.zeros((Q * k + 8,), dtype=torch.int64, device="cuda")
            buf[:Q * k] = keys[g].reshape(-1)
            buf[Q * k:Q * k + 1].view(torch.int32)[0] = flags[g]
            parts.append(buf)
        table = torch.tensor([p.data_ptr() for p in parts], dtype=torch.int64, device="cuda")
        fl = torch.from_numpy(rng.integers(1, 5, size=2 ** 20).astype(np.int32)).cuda()
        qf = torch.from_numpy(rng.integers(1, 5, size=Q).astype(np.int32)).cuda()
        whole = eng.merge_topk(keys, k, q_floor=qf, db_floor_all=fl, max_floor_diff=0, want_keys=True)
        for r in range(G):
            lo, hi = (Q * r) // G, (Q * (r + 1)) // G
            any_flag = torch.full((1,), -7, dtype=torch.int32, device="cuda")
            mine = eng.merge_topk_peers_rows(table.data_ptr(), G, Q, k, lo, hi - lo, q_floor=qf, db_floor_all=fl,
                                             max_floor_diff=0, want_keys=True, flag_offset=Q * k, any_flag=any_flag)
            torch.cuda.synchronize()
            assert torch.equal(mine.keys, whole.keys[lo:hi]) and torch.equal(mine.idx, whole.idx[lo:hi])
            assert torch.equal(mine.scores, whole.scores[lo:hi]) and torch.equal(mine.valid, whole.valid[lo:hi])
            assert torch.equal(mine.count, whole.count[lo:hi])
            assert (int(any_flag.item()) != 0) == any(flags)
            # compaction of a row slice emits global query indices
            oq, om, os_, ov, tot = eng.compact(mine, query_offset=lo)
            wq, wm, ws_, wv, wt = eng.compact(whole)
            t = int(tot.item())
            sel = (wq[:int(wt.item())] >= lo) & (wq[:int(wt.item())] < hi)
            assert t == int(sel.sum()) and torch.equal(oq[:t], wq[:int(wt.item())][sel]) and torch.equal(om[:t], wm[:int(wt.item())][sel])
        # zero rows: only the flags
        any_flag = torch.zeros((1,), dtype=torch.int32, device="cuda")
        eng.merge_topk_peers_rows(table.data_ptr(), G, Q, k, 0, 0, flag_offset=Q * k, any_flag=any_flag)
        assert (int(any_flag.item()) != 0) == any(flags)


@pytest.mark.parametrize("n,d,sym", [(5000, 512, 0), (1500, 128, 1), (9000, 1024, 0)])
def test_one_call_device_sweep_and_graph_replay(eng, n, d, sym):
    """semgate_find_loop_closures_device (K2 + K3 + K4 in one call, caller-owned outputs): eager, captured and replayed
    runs give the candidates of the three-call path, bit for bit; the replays are single graph launches."""
    import torch
    from semgate import _native, synthetic
    desc, ts, fl = synthetic.make_case(n, d, 3, seed=n)
    xb = eng.normalize_cast(_t(desc))
    tts, tfl = _t(ts), _t(fl.astype(np.int32))
    p = _native.make_params(k=25, similarity_threshold=0.5, min_time_gap=10.0, max_floor_diff=0, symmetric=sym,
                            cta_group=2 if sym else 0)
    r = eng.gated_topk(xb, xb, p, q_ts=tts, db_ts=tts, q_floor=tfl, db_floor=tfl)
    wq, wm, ws_, wv, wt = eng.compact(r)
    t = int(wt.item())
    assert t > n // 4
    for use_graph in (False, True):
        for call in range(4):
            oq, om, os_, ov, tot = eng.find_loop_closures_device(xb, p, ts=tts, floor=tfl, use_graph=use_graph)
            torch.cuda.synchronize()
            assert int(tot.item()) == t, (use_graph, call)
            assert torch.equal(oq[:t], wq[:t]) and torch.equal(om[:t], wm[:t]) and torch.equal(os_[:t], ws_[:t]) and torch.equal(ov[:t], wv[:t])
            om.fill_(-5); tot.zero_()
    # a changed database (same shapes and pointers, new contents) is picked up by the replay
    xb[:256].copy_(xb[256:512])
    r2 = eng.gated_topk(xb, xb, p, q_ts=tts, db_ts=tts, q_floor=tfl, db_floor=tfl)
    w2 = eng.compact(r2)
    oq, om, os_, ov, tot = eng.find_loop_closures_device(xb, p, ts=tts, floor=tfl, use_graph=True)
    torch.cuda.synchronize()
    t2 = int(w2[4].item())
    assert int(tot.item()) == t2 and torch.equal(om[:t2], w2[1][:t2]) and torch.equal(os_[:t2], w2[2][:t2])


def test_packed_database_resyncs_every_element(eng):
    """replacing ANY PlaceDescriptor, replacing its descriptor array, or editing a timestamp / floor label
    in place must show in the next call (the reference re-reads everything per call); editing a descriptor's
    contents in place needs invalidate()."""
    from semgate import SemanticPlaceRecognition, PlaceDescriptor, synthetic
    n, d, k = 900, 128, 10
    desc, ts, fl = synthetic.make_case(n, d, 3, seed=12)
    fl32 = fl.astype(np.int32)

    def build():
        spr = SemanticPlaceRecognition('mixvpr', 'cuda', similarity_threshold=0.4, min_time_gap=5.0, descriptor_dim=d)
        for i in range(n):
            spr.vpr.descriptors.append(PlaceDescriptor(float(ts[i]), desc[i].copy(), floor_label=int(fl[i])))
        return spr

    def same(a, b):
        return (np.array_equal(a.query_idx, b.query_idx) and np.array_equal(a.match_idx, b.match_idx)
                and np.array_equal(a.similarity, b.similarity) and np.array_equal(a.is_valid, b.is_valid))

    spr = build()
    base = spr.find_loop_closures_arrays(k=k)
    # (1) replace a record that is neither first, middle nor last
    desc2, ts2, fl2 = desc.copy(), ts.copy(), fl32.copy()
    desc2[137] = desc[400]
    spr.vpr.descriptors[137] = PlaceDescriptor(float(ts[137]), desc2[137].copy(), floor_label=int(fl[137]))
    got = spr.find_loop_closures_arrays(k=k)
    fresh = SemanticPlaceRecognition('mixvpr', 'cuda', similarity_threshold=0.4, min_time_gap=5.0, descriptor_dim=d)
    for i in range(n):
        fresh.vpr.descriptors.append(PlaceDescriptor(float(ts2[i]), desc2[i], floor_label=int(fl2[i])))
    want = fresh.find_loop_closures_arrays(k=k)
    assert same(got, want) and not same(got, base)
    # (2) labels and stamps edited in place; a descriptor array swapped on an existing record
    spr.vpr.descriptors[55].floor_label = 99
    spr.vpr.descriptors[56].timestamp = float(ts[700])
    spr.vpr.descriptors[300].descriptor = desc[301].copy()
    fresh.vpr.descriptors[55].floor_label = 99
    fresh.vpr.descriptors[56].timestamp = float(ts[700])
    fresh.vpr.descriptors[300] = PlaceDescriptor(float(ts2[300]), desc[301].copy(), floor_label=int(fl2[300]))
    assert same(spr.find_loop_closures_arrays(k=k), fresh.find_loop_closures_arrays(k=k))
    # (3) removal from the middle, then contents edited in place + invalidate()
    del spr.vpr.descriptors[10:20]
    del fresh.vpr.descriptors[10:20]
    a, b = spr.find_loop_closures_arrays(k=k), fresh.find_loop_closures_arrays(k=k)
    assert same(a, b) and len(a) > 0
    spr.vpr.descriptors[5].descriptor[:] = spr.vpr.descriptors[600].descriptor
    spr.vpr.invalidate()
    fresh.vpr.descriptors[5] = PlaceDescriptor(fresh.vpr.descriptors[5].timestamp, spr.vpr.descriptors[600].descriptor.copy(),
                                               floor_label=fresh.vpr.descriptors[5].floor_label)
    assert same(spr.find_loop_closures_arrays(k=k), fresh.find_loop_closures_arrays(k=k))
    # (4) emptied and refilled
    spr.vpr.descriptors = []
    assert spr.find_loop_closures(k=k) == []
    spr.vpr.descriptors = list(build().vpr.descriptors)
    assert same(spr.find_loop_closures_arrays(k=k), base)


def test_run_table_is_shared_by_sizes_with_the_same_tile_count(eng):
    """The symmetric run table depends on N only through ceil(N/256): a database growing keyframe by keyframe reuses
    it (the cache was keyed on N and rebuilt + synchronised per call)."""
    from semgate import synthetic
    desc, ts, fl = synthetic.make_case(1300, 128, 3, seed=4)
    fl32 = fl.astype(np.int32)
    for n in (1025, 1026, 1100, 1279, 1280, 1281):
        x, t, f = desc[:n], ts[:n], fl32[:n]         # the same objects on both sides: aliased on the device
        args = (x, x, 15, 0.35, 5.0, t, t, f, f)
        a = run_gpu(eng, *args, mfd=0, cg=2, sym=1)
        b = run_gpu(eng, *args, mfd=0, cg=2, sym=-1)
        assert a["mode"] == 1 and b["mode"] == 0
        assert np.array_equal(a["keys"], b["keys"]), n


def test_clock_probe_reports_the_kernel_clock(eng):
    import torch
    from semgate import _native, synthetic
    x = synthetic.make_descriptors_device(8192, 1024, "cuda", seed=1)
    xb = eng.normalize_cast(x)
    eng.set_option("clock_probe", 1)
    try:
        r = eng.gated_topk(xb, xb, _native.make_params(k=25, similarity_threshold=0.5))
        torch.cuda.synchronize()
        med, mn, span, ctas = eng.clock_probe_read()
    finally:
        eng.set_option("clock_probe", 0)
    assert ctas >= 100 and 300.0 < mn <= med < 2300.0 and span > 10.0
    assert int(r.count.sum()) > 0


def test_full_size_properties_c4(eng, monkeypatch):
    """BASELINE config 4 at full size: AnyLoc-shape 49152-d VLAD descriptors, 250k-keyframe database x 8192 queries
    (24.6 GB of bf16 rows, 768 k-blocks per tile, both operands streaming).  Structure of every list, bit-exact
    decisions on every returned pair, 48 sampled query rows against an fp32 torch reference, and the same lists from
    the other schedule regime (query blocks treated as L2-resident: rm/s chosen differently)."""
    import torch
    from semgate import _native, synthetic
    Q, N, D, F = 8192, 250_000, 49_152, 4
    k, thr, gap = 25, 0.5, 10.0
    free, _ = torch.cuda.mem_get_info()
    if free < 60 * 2 ** 30:
        pytest.skip("needs 60 GB of free device memory")
    dp = _native.pad_dim(D)
    xb = torch.empty((N, dp), dtype=torch.bfloat16, device="cuda")
    g = torch.Generator(device="cuda"); g.manual_seed(11)
    places = max(8, N // 20)
    anchors = torch.randn((places, D), generator=g, device="cuda")
    step = 2048
    for s0 in range(0, N, step):
        e0 = min(N, s0 + step)
        pid = torch.randint(0, places, (e0 - s0,), generator=g, device="cuda")
        eng.normalize_cast(anchors[pid] + 0.6 * torch.randn((e0 - s0, D), generator=g, device="cuda"), out=xb[s0:e0])
    del anchors
    ts = synthetic.make_timestamps(N)
    fl = synthetic.make_floors(N, F).astype(np.int32)
    tts, tfl = _t(ts), _t(fl)
    p = _native.make_params(k=k, similarity_threshold=thr, min_time_gap=gap, max_floor_diff=0)
    qb, qts, qfl = xb[:Q], tts[:Q].contiguous(), tfl[:Q].contiguous()
    shape = _native.schedule_check(Q, N, dp, 2, eng.sm_count)
    assert shape["resident"] == 0, "49152-d query blocks do not fit the L2 budget: both operands stream"
    r = eng.gated_topk(qb, xb, p, q_ts=qts, db_ts=tts, q_floor=qfl, db_floor=tfl, want_keys=True)
    torch.cuda.synchronize()
    monkeypatch.setenv("SEMGATE_RM_CAP_MB", "4096")          # the other regime: "resident" query blocks, rm / s differ
    shape2 = _native.schedule_check(Q, N, dp, 2, eng.sm_count)
    assert (shape2["rm"], shape2["s_main"]) != (shape["rm"], shape["s_main"])
    r2 = eng.gated_topk(qb, xb, p, q_ts=qts, db_ts=tts, q_floor=qfl, db_floor=tfl, want_keys=True)
    torch.cuda.synchronize()
    monkeypatch.delenv("SEMGATE_RM_CAP_MB")
    assert torch.equal(r.keys, r2.keys), "the lists must not depend on the schedule"
    pos = torch.arange(k, device="cuda")[None, :]
    filled = pos < r.count[:, None]
    assert bool(((r.idx >= 0) == filled).all()) and bool((r.scores[filled] >= thr).all())
    assert bool((r.scores[:, 1:][filled[:, 1:]] <= r.scores[:, :-1][filled[:, 1:]]).all()), "scores not descending"
    assert bool((r.idx[filled] < N).all()) and bool((r.valid[~filled] == 0).all())
    qi = torch.arange(Q, device="cuda")[:, None].expand(Q, k)[filled]
    mi = r.idx[filled].long()
    assert not bool(((tts[mi] - tts[qi]).abs() < gap).any()), "a returned pair lies inside the exclusion window"
    assert bool(((tfl[qi] == tfl[mi]) == (r.valid[filled] != 0)).all()), "floor-gate bits differ"
    assert int(r.count.sum().item()) > Q
    rows = np.sort(np.random.default_rng(5).choice(Q, 48, replace=False))
    # fp32 reference on the sampled rows, database converted 4096 rows at a time (fp32 rows are 196 KB each)
    rr = torch.from_numpy(rows).cuda()
    qf32 = qb[rr].float()
    sims = torch.empty((len(rows), N), dtype=torch.float32, device="cuda")
    for s0 in range(0, N, 4096):
        sims[:, s0:s0 + 4096] = qf32 @ xb[s0:s0 + 4096].float().T
    sims[(tts[None, :] - tts[rr][:, None]).abs() < gap] = -float("inf")
    top_s, top_i = torch.topk(sims, k, dim=1)
    top_s, top_i = top_s.cpu().numpy(), top_i.cpu().numpy()
    ref = {"query_idx": [], "match_idx": [], "similarity": [], "is_valid": []}
    for a, row in enumerate(rows):
        for s, j in zip(top_s[a], top_i[a]):
            if np.isfinite(s) and s >= np.float32(thr):
                ref["query_idx"].append(int(row)); ref["match_idx"].append(int(j)); ref["similarity"].append(float(s))
                ref["is_valid"].append(bool(O.floor_ok(fl[row], fl[j], 0)))
    ref = {key: np.asarray(v) for key, v in ref.items()}
    sub = dict(scores=r.scores[rr].cpu().numpy(), idx=r.idx[rr].cpu().numpy().astype(np.int64),
               valid=r.valid[rr].cpu().numpy().astype(bool), count=r.count[rr].cpu().numpy())
    got = O.compact(sub)
    got["query_idx"] = rows[got["query_idx"]]
    rep = parity.compare_candidates(ref, got, k, thr, tol=BF16_MODEL_TOL)
    assert rep["max_score_err"] <= BF16_MODEL_TOL
