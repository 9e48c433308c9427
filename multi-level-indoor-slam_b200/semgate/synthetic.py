"""Seeded synthetic keyframe databases for parity tests and the benchmark.

Shapes follow the descriptor sizes the reference declares for its extractors
(MixVPR 4096-d `place_recognition.py:197`, SALAD 8448-d `:340`, AnyLoc 49152-d
`:418`) and the timestamp / floor-label conventions of its fixtures
(`results/trajectories/lego_loam/5th_floor.txt:1` epoch-second stamps; floor ids
5,1,4,2 from `lego_loam_integration.py:55-60`).

Model: P place anchors shared by every floor (perceptual aliasing: the same
corridor exists on each floor), each keyframe is a noisy copy of one anchor.
Same-place cosine ~0.735, different-place ~0.
"""
from __future__ import annotations

import numpy as np

EPOCH0 = 1678809382.204375  # epoch-scale on purpose: fp32 cannot resolve 10 s here
FLOOR_IDS_3 = (5, 1, 4)
FLOOR_IDS_4 = (5, 1, 4, 2)


def floor_ids(num_floors: int):
    if num_floors == 3:
        return list(FLOOR_IDS_3)
    if num_floors == 4:
        return list(FLOOR_IDS_4)
    return list(range(1, num_floors + 1))


def make_floors(n: int, num_floors: int) -> np.ndarray:
    """Contiguous blocks of equal length, one block per floor (int64 like the
    reference's `np.zeros(n, dtype=int)` labels, `floor_detector.py:134`)."""
    ids = np.asarray(floor_ids(num_floors), dtype=np.int64)
    return ids[(np.arange(n, dtype=np.int64) * num_floors) // max(n, 1)]


def make_timestamps(n: int, dt: float = 0.5, t0: float = EPOCH0) -> np.ndarray:
    return t0 + dt * np.arange(n, dtype=np.float64)


def make_descriptors(n: int, d: int, seed: int = 0, noise: float = 0.6,
                     places: int | None = None, dtype=np.float32) -> np.ndarray:
    """Raw (un-normalised) fp32 descriptors, `[n, d]` C-contiguous."""
    rng = np.random.default_rng(seed)
    p = places if places is not None else max(8, n // 20)
    anchors = rng.standard_normal((p, d), dtype=np.float32)
    pid = rng.integers(0, p, size=n)
    x = anchors[pid]
    x += np.float32(noise) * rng.standard_normal((n, d), dtype=np.float32)
    # arbitrary positive scale per row so that normalisation is actually exercised
    x *= rng.uniform(0.5, 2.0, size=(n, 1)).astype(np.float32)
    return np.ascontiguousarray(x.astype(dtype, copy=False))


def make_case(n: int, d: int, num_floors: int = 3, seed: int = 0, dt: float = 0.5):
    """(descriptors fp32 [n,d], timestamps fp64 [n], floors int64 [n])."""
    return (make_descriptors(n, d, seed), make_timestamps(n, dt), make_floors(n, num_floors))


def make_descriptors_device(n: int, d: int, device, seed: int = 0, noise: float = 0.6,
                            places: int | None = None, chunk: int = 65536):
    """Same model generated on the GPU in chunks (for sizes the host cannot
    hold comfortably).  Returns an fp32 `[n, d]` torch tensor on `device`."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    p = places if places is not None else max(8, n // 20)
    anchors = torch.randn((p, d), generator=g, device=device, dtype=torch.float32)
    out = torch.empty((n, d), device=device, dtype=torch.float32)
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        pid = torch.randint(0, p, (e - s,), generator=g, device=device)
        out[s:e] = anchors[pid]
        out[s:e] += noise * torch.randn((e - s, d), generator=g, device=device, dtype=torch.float32)
    return out


def make_local_features(n: int, patches: int, dim: int, seed: int = 0, places: int | None = None, noise: float = 0.8):
    """Patch-level local features `[n, patches, dim]` fp32 (CricaVPR / DINOv2 style: 529 x 768 at
    322x322 input, place_recognition.py:655,784): every place has its own set of patch anchors,
    a keyframe is a noisy, partly permuted copy (viewpoint change moves patches around)."""
    rng = np.random.default_rng(seed)
    p = places if places is not None else max(2, n // 4)
    anchors = rng.standard_normal((p, patches, dim), dtype=np.float32)
    pid = rng.integers(0, p, size=n)
    out = np.empty((n, patches, dim), dtype=np.float32)
    for i in range(n):
        perm = np.arange(patches)
        sw = rng.integers(0, patches, size=(patches // 4, 2))
        perm[sw[:, 0]], perm[sw[:, 1]] = perm[sw[:, 1]], perm[sw[:, 0]]
        out[i] = anchors[pid[i]][perm] + np.float32(noise) * rng.standard_normal((patches, dim), dtype=np.float32)
        out[i] *= rng.uniform(0.5, 3.0, size=(patches, 1)).astype(np.float32)
    return out, pid
