"""CPU model of the fold K3's dense-list kernel uses (csrc/kernels.cu, merge_dense_kernel): a running list of W slots,
sorted descending with the slots beyond k zero, takes another such list by one max-against-the-reverse step (the result
holds the W largest of both and is bitonic) and a bitonic merge; only the top k go on.  Checked against a plain sort for
every k <= 32 at the widths the kernel instantiates (8 / 16 / 32), with ties, short and empty lists."""
import numpy as np
import pytest


def bitonic_merge_desc(c, w):
    c = c.copy()
    stride = w // 2
    while stride >= 1:
        for i in range(w):
            if (i & stride) == 0:
                a, b = c[:, i].copy(), c[:, i | stride].copy()
                c[:, i], c[:, i | stride] = np.maximum(a, b), np.minimum(a, b)
        stride //= 2
    return c


def fold(best, cur, w, k):
    out = np.maximum(best[:, :w], cur[:, :w][:, ::-1])
    out = bitonic_merge_desc(out, w)
    out[:, k:] = 0
    return out


@pytest.mark.parametrize("k", list(range(1, 33)))
def test_fold_equals_sort(k):
    w = 8 if k <= 8 else 16 if k <= 16 else 32
    rng = np.random.default_rng(k)
    rows, lists = 400, 5
    keys = rng.integers(1, 2 ** 62, size=(lists, rows, w), dtype=np.int64).astype(np.uint64)
    keys[:, :, k:] = 0
    fill = rng.integers(0, k + 1, size=(lists, rows))
    keys[np.arange(w)[None, None, :] >= fill[:, :, None]] = 0
    keys[1, ::7] = keys[0, ::7]                      # exact ties between lists
    keys = np.sort(keys, axis=2)[:, :, ::-1]         # every list sorted descending, zeros last
    best = keys[0].copy()
    for g in range(1, lists):
        best = fold(best, keys[g], w, k)
    want = np.sort(keys.transpose(1, 0, 2).reshape(rows, lists * w), axis=1)[:, ::-1][:, :k]
    assert np.array_equal(best[:, :k], want)
    assert not best[:, k:].any()
