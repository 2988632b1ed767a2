"""GPU checks of the `ami` mirror (reference measures/mi.py:212-262, EfficientAMI) through get_measure('ami')."""
import os

import numpy as np
import pytest
import torch

from oracle import batch_mi_oracle as bo, gen_golden

pytestmark = pytest.mark.gpu

AMI = sorted(gen_golden.AMI_CASES)


def gpu_measure(a, C, **kw):
    from acav100m_b200.subset_selection import get_measure
    return get_measure("ami")(a, ncentroids=C, device="cuda", **kw)


@pytest.mark.parametrize("name", AMI)
def test_scores_follow_the_reference_iteration_by_iteration(golden_dir, name):
    """Teacher-forced replay of the reference's own `ami` run: every score of every remaining candidate within 1e-5
    of the reference's fp32 value and within 1e-6 of the fp64 evaluation of the same expressions; the same pick
    whenever the best two DIFFERENT score values are further apart than the reference's fp32 noise."""
    g = dict(np.load(os.path.join(golden_dir, name + ".npz")))
    a = g["assignments"].astype(np.int64)
    order = g["candidate_order"].tolist()
    C, pairs, subset = int(g["c"]), [tuple(p) for p in g["pairs"].tolist()], int(g["subset"])
    S_ref = g["S"].tolist()
    S, GAIN, ALL = bo.greedy_dense_mi(a, C, pairs, order[1:], subset, [order[0]], follow=S_ref[1:], measure="ami")
    assert S == S_ref and np.array_equal(np.array(GAIN), g["GAIN"])          # the oracle replay IS the golden run
    _, _, ALL64 = bo.greedy_dense_mi(a, C, pairs, order[1:], subset, [order[0]], follow=S_ref[1:], measure="ami",
                                     dtype=torch.float64)
    m = gpu_measure(a, C)
    m.init(pairs, order[1:])
    decided = 0
    for it, ((want, cand), (want64, _)) in enumerate(zip(ALL, ALL64)):
        assert torch.equal(m.candidate_ids, cand)
        got = m.score_candidates().cpu()
        np.testing.assert_allclose(got.numpy(), want.numpy(), rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(got.numpy(), want64.numpy(), rtol=1e-6, atol=1e-7)
        top = want.max().item()
        near = want[(want >= top - 1e-5 * max(abs(top), 1e-3))]
        if torch.unique(near).numel() == 1:                        # no near-tie between different values
            assert int(got.max(dim=0).indices) == int(want.max(dim=0).indices)
            decided += 1
        idx = int((cand == S_ref[1 + it]).nonzero()[0, 0])         # teacher forcing: follow the reference's pick
        m._add_cells(m._cand_cells[idx:idx + 1].contiguous())
        m.remove_idx_all(idx)
    # the reference's run sits on a plateau of AMI = 1 +- 1e-7 picks (ties between different cells decided by its fp32
    # noise), so only part of the iterations have a clear winner: 11 of 58 for ami_small, 36 of 38 for ami_p3
    assert decided >= 10


@pytest.mark.parametrize("method", ["arithmetic", "max", "min"])
def test_average_methods_and_larger_tables(method):
    """generalized_mean variants (mi.py:200-209) at a size the dense fp64 oracle still evaluates: C = 24, three pairs,
    2500 candidates scored against a table holding 600 samples."""
    rng = np.random.RandomState(12)
    C, V = 24, 3100
    base = rng.randint(0, C, size=(V, 1))
    a = np.where(rng.random_sample((V, 3)) < 0.6, (base * 5 + np.arange(3)) % C, rng.randint(0, C, size=(V, 3))).astype(np.int64)
    pairs = [(0, 1), (0, 2), (1, 2)]
    m = gpu_measure(a, C, average_method=method)
    cands = list(range(600, V))
    m.init(pairs, cands)
    m.add_samples(list(range(600)))
    got = m.score_candidates().cpu().numpy()
    # dense fp64 evaluation of (table of the first 600 rows + candidate)
    at = torch.from_numpy(a)
    N = torch.full((3, C, C), bo.EPS, dtype=torch.float64)
    cache = {"N": N, "a": N.sum(dim=1), "b": N.sum(dim=2)}
    cache["n"] = cache["a"].sum(dim=-1)
    add = bo.sample_tables(at, pairs, torch.arange(600), C)
    cache = {k: cache[k] + add[k].to(torch.float64).sum(0) for k in cache}
    want = []
    for lo in range(600, V, 500):
        tabs = bo.sample_tables(at, pairs, torch.arange(lo, min(lo + 500, V)), C)
        last = {k: cache[k].unsqueeze(0) + tabs[k].to(torch.float64) for k in tabs}
        want.append(bo.dense_ami(last, method).mean(dim=-1))
    want = torch.cat(want).numpy()
    np.testing.assert_allclose(got, want, rtol=2e-6, atol=1e-7)


def test_free_running_selection_and_registry(golden_dir):
    from acav100m_b200.subset_selection import get_measure
    from acav100m_b200.subset_selection.measures import EfficientAMI
    assert get_measure("AMI") is EfficientAMI
    g = dict(np.load(os.path.join(golden_dir, "ami_small.npz")))
    a = g["assignments"].astype(np.int64)
    order = g["candidate_order"].tolist()
    m = gpu_measure(a, int(g["c"]))
    m.init([tuple(p) for p in g["pairs"].tolist()], order[1:])
    S, GAIN, timelapse, LOOKUPS = m.run_greedy(int(g["subset"]), [order[0]])
    assert len(S) == int(g["subset"]) - 1 and len(set(S)) == len(S)
    assert len(GAIN) == len(timelapse) == len(LOOKUPS) == len(S) - 1 and all(np.isfinite(GAIN))
    # a free run leaves the reference's trajectory at the first tie between different cells (decided by fp32 noise in
    # the reference; its own fp64 evaluation free-runs to different picks too), so beyond the invariants only the first
    # two scores are comparable: an empty table scores 0, one sample on its own diagonal cell scores 1
    np.testing.assert_allclose(GAIN[:2], g["GAIN"][:2], rtol=1e-5, atol=1e-6)
