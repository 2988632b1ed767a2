"""GPU checks of the dense `mi` mirror (reference measures/mi.py:14-209, the toy-size cross-check measure)."""
import os

import numpy as np
import pytest
import torch

from oracle import batch_mi_oracle as bo

pytestmark = pytest.mark.gpu


def gpu_measure(a, C, exact=True):
    from acav100m_b200.subset_selection import get_measure
    return get_measure("mi")(a, ncentroids=C, device="cuda", exact=exact)


@pytest.mark.parametrize("exact", [True, False])
def test_scores_follow_the_reference_iteration_by_iteration(golden_dir, exact):
    """Teacher-forced replay of the reference's own `mi` run (tests/golden/mi_dense_small_mi.npz): every score of
    every remaining candidate within 1e-5 relative, and the same pick whenever the best two DIFFERENT score
    values are further apart than that."""
    g = dict(np.load(os.path.join(golden_dir, "mi_dense_small_mi.npz")))
    a = g["assignments"].astype(np.int64)
    order = g["candidate_order"].tolist()
    C, pairs, subset = int(g["c"]), [tuple(p) for p in g["pairs"].tolist()], int(g["subset"])
    S_ref = g["S"].tolist()
    S, GAIN, ALL = bo.greedy_dense_mi(a, C, pairs, order[1:], subset, [order[0]], follow=S_ref[1:])
    assert S == S_ref                                              # the oracle replay reproduces the golden picks
    np.testing.assert_allclose(GAIN, g["GAIN"], rtol=1e-5, atol=1e-9)
    m = gpu_measure(a, C, exact)
    m.init(pairs, order[1:])
    decided = 0
    for it, (want, cand) in enumerate(ALL):
        assert torch.equal(m.candidate_ids, cand)
        got = m.score_candidates().cpu()
        np.testing.assert_allclose(got.numpy(), want.numpy(), rtol=1e-5, atol=1e-7)
        if exact:                                                  # the reference's bits, hence its arg-max
            assert np.array_equal(got.numpy(), want.numpy())
            assert int(got.max(dim=0).indices) == int(want.max(dim=0).indices)
        top = want.max().item()
        others = want[want < top - 1e-5 * max(abs(top), 1e-3)]
        near = want[(want >= top - 1e-5 * max(abs(top), 1e-3))]
        if torch.unique(near).numel() == 1:                        # no near-tie between different values
            assert int(got.max(dim=0).indices) == int(want.max(dim=0).indices)
            decided += 1
        idx = int((cand == S_ref[1 + it]).nonzero()[0, 0])         # teacher forcing: follow the reference's pick
        m._add_cells(m._cand_cells[idx:idx + 1].contiguous())
        m.remove_idx_all(idx)
        del others
    assert decided > len(ALL) // 2


def test_free_running_selection(golden_dir):
    g = dict(np.load(os.path.join(golden_dir, "mi_dense_small_mi.npz")))
    a = g["assignments"].astype(np.int64)
    order = g["candidate_order"].tolist()
    m = gpu_measure(a, int(g["c"]))
    m.init([tuple(p) for p in g["pairs"].tolist()], order[1:])
    S, GAIN, timelapse, LOOKUPS = m.run_greedy(int(g["subset"]), [order[0]])
    assert len(S) == int(g["subset"]) - 1 and len(set(S)) == len(S)
    assert len(GAIN) == len(timelapse) == len(LOOKUPS) == len(S) - 1
    # exact scoring (default): index for index the reference's CPU run, with its fp32 scores
    assert S == g["S"].tolist()
    assert np.array_equal(np.array(GAIN, dtype=np.float32), g["GAIN"].astype(np.float32))


def test_three_pairs_and_errors():
    rng = np.random.RandomState(3)
    a = rng.randint(0, 5, size=(120, 3)).astype(np.int64)
    pairs = [(0, 1), (0, 2), (1, 2)]
    S, GAIN, ALL = bo.greedy_dense_mi(a, 5, pairs, list(range(1, 120)), 30, [0])
    m = gpu_measure(a, 5)
    m.init(pairs, list(range(1, 120)))
    got = m.score_candidates().cpu()
    np.testing.assert_allclose(got.numpy(), ALL[0][0].numpy(), rtol=1e-5, atol=1e-7)
    S2, GAIN2, _, _ = m.run_greedy(30, [0])
    # the very first pick is a tie of all candidates at MI = 0 +- 1e-13 which the reference settles by the rounding
    # noise of its dense sum: with the exact scorer the free run follows the oracle's trajectory index for index
    assert S2 == S and np.array_equal(np.array(GAIN2, dtype=np.float32), np.array(GAIN, dtype=np.float32))
    assert len(S2) == 29 and len(set(S2)) == 29 and all(np.isfinite(GAIN2))
    bad = gpu_measure(a, 4)
    with pytest.raises(ValueError):
        bad.init(pairs, list(range(1, 120)))
