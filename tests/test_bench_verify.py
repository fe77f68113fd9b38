"""bench.py's self-verification (the `"verified"` key of the bench line) is itself checked here, on
the CPU, against the oracle: it must accept the oracle's own stable top-k, tolerate a swap inside a
near-tie, and reject a wrong index or a wrong score."""
from collections import namedtuple

import torch

import bench
from oracle import reference_path as ref
from oracle import synthetic

TopK = namedtuple("TopK", "values indices")


def _setup(Q=64, N=3000, D=64, k=3):
    q, g = synthetic.iid_gallery(N, D, Q, seed=12, dtype=torch.bfloat16)
    v, i, _ = ref.cos_topk_stable(q, g, k)
    return q, g, TopK(v.float(), i.clone())


def _verify(q, g, res, k=3):
    return bench.verify_against_torch(None, None, lambda _q: res, q, g, 0, g.shape[0], k, 1,
                                      torch.device("cpu"))


def test_accepts_the_oracle_result():
    q, g, res = _setup()
    out = _verify(q, g, res)
    assert out["verified"] and out["bad_indices"] == 0 and out["queries_checked"] == 16
    assert out["max_abs_score_err"] < 1e-5 and out["indices_identical_frac"] == 1.0


def test_rejects_a_wrong_index_and_a_wrong_score():
    q, g, res = _setup()
    bad = TopK(res.values.clone(), res.indices.clone())
    bad.indices[0, 0] = (bad.indices[0, 0] + 1234) % g.shape[0]       # row 0 is one of the sampled queries
    assert not _verify(q, g, bad)["verified"]
    off = TopK(res.values.clone(), res.indices.clone())
    off.values[0, 1] += 0.05
    assert not _verify(q, g, off)["verified"]


def test_tolerates_a_swap_inside_an_exact_tie():
    q, g, _ = _setup()
    g[17] = g[5]                                                       # exact duplicate rows
    q[0] = g[5] * 2                                                    # query 0's best match is the tie
    v, i, _ = ref.cos_topk_stable(q, g, 3)
    assert set(i[0, :2].tolist()) == {5, 17}
    swapped = TopK(v.float(), i.clone())
    swapped.indices[0, 0], swapped.indices[0, 1] = i[0, 1].clone(), i[0, 0].clone()
    # torch.topk's own order inside a tie is unspecified: either order of the tied rows verifies
    for res in (TopK(v.float(), i.clone()), swapped):
        out = _verify(q, g, res)
        assert out["verified"] and out["bad_indices"] == 0
