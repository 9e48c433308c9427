"""CPU tier: the C-ABI library builds, loads without a GPU, and exports every symbol
include/semgate.h declares.  No compute calls here."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "semgate.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(semgate_[a-z_0-9]+)\s*\(", src)))


def test_library_builds_and_exports_header_symbols():
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "semgate_build", os.path.join(ROOT, "multi-level-indoor-slam_b200", "build.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    path = b.build()
    lib = ctypes.CDLL(path)
    declared = _declared_symbols()
    assert len(declared) >= 18
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/semgate.h but not exported"
    from semgate import _native
    assert sorted(_native.SYMBOLS) == declared, "ctypes binding and header disagree"
    lib.semgate_version.restype = ctypes.c_int
    assert lib.semgate_version() == 200
    assert lib.semgate_pad_dim(4096) == 4096 and lib.semgate_pad_dim(100) == 128 and lib.semgate_pad_dim(8448) == 8448


def test_params_struct_layout():
    from semgate import _native
    # float, (pad), double, 9 x 32-bit -> 52, rounded to the double's alignment
    assert ctypes.sizeof(_native.TopkParams) == 56
    assert _native.TopkParams.symmetric.offset == 40 and _native.TopkParams.part_count.offset == 48
    assert _native.TopkParams.min_time_gap.offset == 8 and _native.TopkParams.k.offset == 16
    p = _native.make_params(k=25, similarity_threshold=0.5, min_time_gap=10.0, max_floor_diff=0)
    assert p.k == 25 and p.similarity_threshold == 0.5 and p.gate_mode == 0


@pytest.mark.parametrize("sym", [False, True])
def test_tile_schedule_invariants(sym):
    """The schedule every warp role of the fused kernel walks, checked on the host: each tile computed exactly
    once (symmetric sweep: exactly the tiles on or above the block diagonal), list slots unique per block,
    pacing counters in range and arrivals equal to what the waiters expect."""
    import numpy as np
    from semgate import _native
    rng = np.random.default_rng(7)
    sizes = [257, 300, 512, 513, 1000, 4096, 5000, 18945, 20000, 100000, 300000, 1000000]
    sizes += [int(v) for v in rng.integers(257, 60000, size=24)]
    for n in sizes:
        for d_pad in (64, 512, 4096, 8448, 49152):
            if n > 100000 and d_pad != 4096:
                continue
            for sms in (148, 132, 16):
                if n > 100000 and sms != 148:
                    continue
                r = _native.schedule_check(n, n, d_pad, 2, sms, sym)
                nb = (n + 255) // 256
                assert r["blocks"] == nb and r["tiles"] == nb
                assert r["computed"] == (nb * (nb + 1) // 2 if sym else nb * nb)
                if sym and nb <= 224:      # the run table such sweeps actually use
                    t = _native.schedule_check(n, n, d_pad, 2, sms, 2)
                    assert t["computed"] == nb * (nb + 1) // 2 and t["makespan"] <= r["makespan"]
    if not sym:   # rectangular problems, single-CTA tiles
        for Q, N, cg in [(1, 1, 1), (100, 5000, 1), (10000, 100000, 2), (8192, 250000, 2), (1500, 300, 1), (129, 257, 2)]:
            r = _native.schedule_check(Q, N, 4096, cg)
            assert r["computed"] == r["blocks"] * r["tiles"]
    else:
        full = _native.schedule_check(1000000, 1000000, 4096, 2, 148, False)
        half = _native.schedule_check(1000000, 1000000, 4096, 2, 148, True)
        assert half["makespan"] < 0.51 * full["makespan"]
        with pytest.raises(_native.SemgateError):
            _native.schedule_check(1000, 2000, 4096, 2, 148, True)
        # split over G GPUs: the parts tile the triangle exactly once and are balanced
        for n, G, how in [(20000, 2, 1), (56568, 8, 1), (300000, 4, 1), (1000000, 8, 1), (5000, 3, 1),
                          (28160, 2, 2), (39936, 4, 2), (56320, 8, 2), (700, 8, 2)]:
            parts = [_native.schedule_check(n, n, 4096, 2, 148, how, g, G) for g in range(G)]
            nb = (n + 255) // 256
            assert sum(p["computed"] for p in parts) == nb * (nb + 1) // 2
            if n >= 300000:
                assert max(p["makespan"] for p in parts) <= 1.03 * sum(p["makespan"] for p in parts) / G
        with pytest.raises(_native.SemgateError):
            _native.schedule_check(5000, 5000, 4096, 2, 148, False, 1, 2)      # only symmetric sweeps split this way


def test_torch_ops_register_without_a_gpu():
    """torch.ops.semgate.* (the thin PyTorch extension over the C ABI): builds, loads, registers its four
    operators with CUDA kernels only -- a CPU tensor has nothing to dispatch to."""
    import importlib.util
    import torch
    spec = importlib.util.spec_from_file_location(
        "semgate_build", os.path.join(ROOT, "multi-level-indoor-slam_b200", "build.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    b.build_torch_ops()
    from semgate import ops
    sg = ops.load()
    for name in ("normalize_cast", "gated_topk", "merge_topk", "compact"):
        assert hasattr(sg, name)
    schema = str(torch.ops.semgate.gated_topk.default._schema)
    assert "Tensor? q_floor" in schema and "int db_index_offset" in schema
    with pytest.raises((NotImplementedError, RuntimeError)):
        sg.normalize_cast(torch.zeros(4, 64))


def test_no_gpu_fails_loudly():
    """Without a usable sm_100 device the product path raises; it never falls back."""
    import numpy as np
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from semgate import _native, SemanticPlaceRecognition, PlaceDescriptor
    with pytest.raises(_native.SemgateError):
        _native.Engine(0)
    spr = SemanticPlaceRecognition('mixvpr', 'cuda', descriptor_dim=8)
    for i in range(3):
        spr.vpr.descriptors.append(PlaceDescriptor(float(i), np.ones(8, np.float32), floor_label=1))
    with pytest.raises(Exception):
        spr.find_loop_closures()
    with pytest.raises(RuntimeError):
        SemanticPlaceRecognition('mixvpr', 'cpu', descriptor_dim=8).vpr._engine()


def test_interface_parity_with_reference_signatures():
    """Same constructor / method signatures as the reference classes
    (place_recognition.py:61-190, :806-933; loop_closure_gate.py:16-148)."""
    import inspect
    import semgate

    def sig(f):
        return list(inspect.signature(f).parameters)

    assert sig(semgate.PlaceMatch)[:6] == ["query_idx", "match_idx", "similarity", "query_timestamp",
                                           "match_timestamp", "is_valid"]
    assert sig(semgate.PlaceDescriptor) == ["timestamp", "descriptor", "image_path", "floor_label"]
    assert sig(semgate.BasePlaceRecognition.add_image) == ["self", "image", "timestamp", "floor_label", "image_path"]
    assert sig(semgate.BasePlaceRecognition.query) == ["self", "image", "timestamp", "k", "min_time_gap"]
    assert sig(semgate.SemanticPlaceRecognition.__init__)[:5] == ["self", "vpr_method", "device",
                                                                  "similarity_threshold", "min_time_gap"]
    assert sig(semgate.SemanticPlaceRecognition.find_loop_closures)[:3] == ["self", "enable_floor_gating", "k"]
    assert sig(semgate.SemanticLoopClosureGate.__init__)[:3] == ["self", "floor_labels", "strict_mode"]
    assert sig(semgate.SemanticLoopClosureGate.gate_candidate) == ["self", "query_idx", "match_idx", "similarity_score"]
    assert sig(semgate.LoopClosureCandidate) == ["query_idx", "match_idx", "similarity_score", "query_floor",
                                                 "match_floor", "is_valid", "rejection_reason"]
    d = inspect.signature(semgate.SemanticPlaceRecognition.__init__).parameters
    assert d["similarity_threshold"].default == 0.5 and d["min_time_gap"].default == 10.0
    assert inspect.signature(semgate.SemanticPlaceRecognition.find_loop_closures).parameters["k"].default == 10
    assert inspect.signature(semgate.BasePlaceRecognition.query).parameters["k"].default == 5
    spr = semgate.SemanticPlaceRecognition('salad')
    assert spr.vpr.descriptor_dim == 8448
    assert spr.get_statistics([]) == {'total_matches': 0, 'valid_matches': 0, 'rejected_matches': 0,
                                      'rejection_rate': 0.0}
    with pytest.raises(ValueError):
        semgate.SemanticPlaceRecognition(vpr_method="nope")
    g = semgate.SemanticLoopClosureGate([1, 2])
    assert g.gate_candidates([]) == ([], []) and g.get_stats()["total_candidates"] == 0


def test_cricavpr_host_logic_without_gpu():
    """Feature caching and the no-rerank short cuts never touch the device (reference :733-737, :768-777)."""
    import numpy as np
    import semgate
    v = semgate.CricaVPR(device='cuda', use_reranking=True)
    lf = np.ones((7, 16), np.float32)
    v.add_image(np.zeros(v.descriptor_dim, np.float32), 0.0, 1, local_features=lf)
    v.add_image(np.zeros(v.descriptor_dim, np.float32), 1.0, 1)
    assert list(v._feature_cache) == [0] and v._feature_cache[0].shape == (1, 7, 16)
    assert v.extract_local_features(lf[None]).shape == (1, 7, 16)
    with pytest.raises(NotImplementedError):
        v.extract_local_features(np.zeros((2, 3, 4, 5)))
    cands = [(0, 0.9), (1, 0.8), (2, 0.7)]
    assert v.rerank_candidates(5, cands, top_k=2) == cands[:2]          # query without cached features
    v.use_reranking = False
    assert v.rerank_candidates(0, cands, top_k=5) == cands
    w = semgate.CricaVPR(use_reranking=False)
    w.add_image(np.zeros(w.descriptor_dim, np.float32), 0.0, 1, local_features=lf)
    assert w._feature_cache == {}
