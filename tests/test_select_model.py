"""Host model of the histogram form of the large-k selection (csrc/topk_select.cu, rows of up to
32768 scores): the bin function, the search for the k-th best's bin and the collect-and-sort step
restated in numpy and checked against a full stable sort — including the inputs the GPU tests
reach only by accident: NaN / +-Inf scores, scores outside [-1, 1], rows shorter than k, and rows
whose k-th best sits in a bin too crowded for the 1024-key buffer (where the kernel must fall back
to the streaming form).  Reference semantics: torch.topk(sim, k=150) of
inference/training_analysis.ipynb:238, NaN first, ties -> lower gallery index.
"""
import numpy as np
import pytest

SEL_BINS, SEL_KEYS = 2048, 1024


def score_bin(v):
    """score_bin(): linear bins of 1/1024 over [-1, 1], clamped; NaN in the top bin."""
    v = np.asarray(v, dtype=np.float32)
    with np.errstate(invalid="ignore", over="ignore"):
        x = (v + np.float32(1.0)) * np.float32(SEL_BINS // 2)
        x = np.minimum(np.maximum(x, np.float32(0.0)), np.float32(SEL_BINS - 1))
    b = np.where(np.isnan(v), SEL_BINS - 1, np.nan_to_num(x, nan=0.0)).astype(np.int64)
    return b


def rank_key(v, idx):
    """Descending rank order used everywhere: NaN first, then score descending, then index ascending."""
    nan_first = np.where(np.isnan(v), 0, 1)
    vv = np.where(np.isnan(v), 0.0, v)
    return np.lexsort((idx, -vv, nan_first))


def select_hist(row, k):
    """The kernel's histogram form; returns (indices, used_fallback)."""
    n = row.shape[0]
    bins = score_bin(row)
    hist = np.bincount(bins, minlength=SEL_BINS)
    suffix = np.cumsum(hist[::-1])[::-1]               # scores in bins >= b
    ok = np.nonzero(suffix >= k)[0]
    kth_bin = int(ok.max()) if ok.size else 0           # fewer than k scores: take them all
    cand = np.nonzero(bins >= kth_bin)[0]
    if cand.size > SEL_KEYS:
        return None, True                               # crowded bin -> streaming form
    order = cand[rank_key(row[cand], cand)]
    return order[:k], False


def oracle(row, k):
    idx = np.arange(row.shape[0])
    return idx[rank_key(row, idx)][:k]


def test_bin_function_is_monotone_in_the_ranking_order():
    rng = np.random.default_rng(0)
    v = np.concatenate([rng.standard_normal(5000).astype(np.float32) * 0.3,
                        np.float32([-np.inf, np.inf, -1.0, 1.0, -1.0000001, 1.0000001, 0.0, -0.0, 5.0, -7.0]),
                        np.float32([np.nan, np.nan])])
    idx = np.arange(v.shape[0])
    order = idx[rank_key(v, idx)]                        # best first
    b = score_bin(v)[order]
    assert (b[:-1] >= b[1:]).all()
    assert b.min() >= 0 and b.max() <= SEL_BINS - 1


@pytest.mark.parametrize("n,k", [(8736, 150), (8736, 256), (300, 17), (100, 100), (40, 256), (32768, 150)])
def test_histogram_selection_equals_a_full_sort(n, k):
    rng = np.random.default_rng(n + k)
    for trial in range(6):
        row = (rng.standard_normal(n) * (0.02 if trial % 2 else 0.2)).astype(np.float32)
        if trial >= 2:                                   # duplicates of the best scores, NaN, Inf
            row[rng.integers(0, n, n // 10)] = row[rng.integers(0, n, 1)]
        if trial >= 4 and n > 50:
            row[rng.integers(0, n, 3)] = np.nan
            row[rng.integers(0, n, 2)] = np.inf
            row[rng.integers(0, n, 2)] = -np.inf
        got, fell_back = select_hist(row, k)
        assert not fell_back
        want = oracle(row, k)
        assert np.array_equal(got, want[: got.shape[0]]) and got.shape[0] == min(k, n)


def test_a_crowded_bin_is_detected():
    row = np.full(6000, 0.25, dtype=np.float32)
    row[:10] = 0.9
    got, fell_back = select_hist(row, 150)
    assert fell_back                                     # 5990 equal scores share the k-th best's bin
    got, fell_back = select_hist(row, 5)                 # the k-th best is above the crowd
    assert not fell_back and np.array_equal(got, np.arange(5))
