"""CPU tier: the fp32 pre-test of the temporal-exclusion window (gated_topk.cuh::window_excluded) emulated in numpy
float32 (IEEE round-to-nearest, like the device) against the reference's fp64 predicate `abs(t_db - t_q) < gap`
(place_recognition.py:884): the two "definite" branches must never disagree with fp64, whatever the stamps."""
import numpy as np
import pytest

f32 = np.float32


def gap_neighbours(gap):
    g = f32(gap)
    lo = np.nextafter(g, f32(-np.inf)) if float(g) > gap else g
    hi = np.nextafter(g, f32(np.inf)) if float(g) < gap else g
    return f32(lo), f32(hi)


def pretest(t_db, t_q, base, gap):
    """-> (decided: bool array, excluded where decided)"""
    a = (t_db - base).astype(f32)            # fl32(fl64(t - base))
    b = (t_q - base).astype(f32)
    d = np.abs(a - b)                        # fp32 subtraction
    m = (np.abs(a) + np.abs(b) + d) * f32(2.0 ** -22) + f32(1e-30)
    lo, hi = gap_neighbours(gap)
    with np.errstate(invalid="ignore"):
        not_excl = d > hi + m
        excl = d < lo - m
    return not_excl | excl, excl


@pytest.mark.parametrize("seed", range(8))
def test_pretest_never_contradicts_fp64(seed):
    rng = np.random.default_rng(seed)
    n = 200000
    base = 1678809382.204375 if seed % 2 == 0 else float(rng.uniform(-1e6, 1e6))
    span = float(rng.choice([50.0, 5e3, 5e5, 5e7]))
    t_q = base + rng.uniform(-span, span, n)
    gap = float(rng.choice([0.0, 0.5, 10.0, 10.000000001, 3600.0, 1e-9, -1.0]))
    # adversarial: most pairs sit within a few fp32/fp64 ulps of the window edge
    off = gap + rng.choice([0.0, 1.0, -1.0, 3.0, -3.0], n) * np.spacing(f32(max(abs(gap), 1e-30))).astype(np.float64) \
        * rng.uniform(0, 4, n)
    far = rng.random(n) < 0.3
    off[far] = rng.uniform(0, 2 * span + 1, int(far.sum()))
    t_db = t_q + rng.choice([-1.0, 1.0], n) * off
    want = np.abs(t_db - t_q) < gap                      # the reference predicate, fp64
    decided, excl = pretest(t_db, t_q, base, gap)
    assert np.array_equal(excl[decided], want[decided])
    if gap > 0 and span <= 5e5:
        assert decided[far].mean() > 0.99                # the exact path is the rare one


def test_pretest_special_values():
    base = 1000.0
    t_q = np.array([1000.0, np.nan, np.inf, 5.0, 5.0])
    t_db = np.array([np.nan, 3.0, 7.0, -np.inf, 5.0])
    for gap in (0.0, 10.0, np.inf):
        decided, excl = pretest(t_db, t_q, base, gap)
        with np.errstate(invalid="ignore"):
            want = np.abs(t_db - t_q) < gap
        assert np.array_equal(excl[decided], want[decided])
    decided, _ = pretest(t_db[:4], t_q[:4], base, 10.0)
    assert not decided.any()                             # NaN / inf stamps always take the exact test
