"""GPU checks of the `batch_mi` mirror (the reference CLI's default measure)."""
import os

import numpy as np
import pytest
import torch

from oracle import batch_mi_oracle as bo, gen_golden

pytestmark = pytest.mark.gpu

CASES = sorted(gen_golden.BATCH_MI_CASES)


def load(golden_dir, name):
    return dict(np.load(os.path.join(golden_dir, name + ".npz")))


def gpu_measure(a, C, keep_unselected=True, exact=True):
    from acav100m_b200.subset_selection import get_measure
    return get_measure("batch_mi")(a, ncentroids=C, batch_size=min(20, a.shape[0] - 1), selection_size=4,
                                   device="cuda", keep_unselected=keep_unselected, exact=exact)


@pytest.mark.parametrize("exact", [True, False])
@pytest.mark.parametrize("name", CASES)
def test_scores_and_topk_follow_the_oracle_iteration_by_iteration(golden_dir, name, exact):
    """Teacher-forced replay: feed the oracle's own batches and picks to the CUDA engine.  exact=True
    (acav_mi_dense_score_exact, the default): every score carries the reference's BITS and topk returns the same
    indices in the same order.  exact=False (O(B*P) from fp64 running sums): every score within 1e-5 relative and the
    top-k SET equal whenever it is not decided by a near-tie."""
    g = load(golden_dir, name)
    a = g["assignments"].astype(np.int64)
    C, pairs = int(g["c"]), [tuple(p) for p in g["pairs"].tolist()]
    keep = bool(g["keep_unselected"])
    V, subset = a.shape[0], int(g["subset"])
    m = gpu_measure(a, C, keep, exact)
    m.init(pairs, list(range(1, V)))
    m.k = m.modify_k(subset)
    m.add_samples([0])
    # oracle state, stepped with the reference's arithmetic
    at = torch.from_numpy(a)
    N = torch.full((len(pairs), C, C), bo.EPS)
    cache = {"N": N, "a": N.sum(1), "b": N.sum(2)}
    cache["n"] = cache["a"].sum(-1)
    add = bo.sample_tables(at, pairs, torch.tensor([0]), C)
    for key in cache:
        cache[key] = cache[key] + add[key].sum(0)
    cand = torch.arange(1, V)
    gen_golden.seed_all(int(g["seed"]))
    picked, decided, undecided = [], 0, 0
    k = bo.modify_k(4, m.B, subset, V, keep)
    while len(picked) < subset:
        cand = cand.index_select(0, torch.randperm(cand.shape[0]))
        batch = cand[:m.B]
        tabs = bo.sample_tables(at, pairs, batch, C)
        want = bo.dense_mi({key: cache[key].unsqueeze(0) + tabs[key] for key in tabs}).mean(-1)
        got = m.score_batch(batch)
        np.testing.assert_allclose(got.numpy(), want.numpy(), rtol=1e-5, atol=1e-7)
        top, ids = want.topk(k)
        if exact:
            assert np.array_equal(got.numpy(), want.numpy()), "scores must carry the reference's bits"
            assert got.topk(k).indices.tolist() == ids.tolist()
        srt = want.sort(descending=True).values
        gap = (srt[k - 1] - srt[k]).item() if len(srt) > k else 1.0
        if gap > 1e-5 * max(abs(srt[k - 1].item()), 1e-3):
            assert set(got.topk(k).indices.tolist()) == set(ids.tolist())
            decided += 1
        else:
            undecided += 1
        chosen = batch.index_select(0, ids)                       # teacher forcing: follow the oracle's picks
        win = bo.sample_tables(at, pairs, chosen, C)
        for key in cache:
            cache[key] = cache[key] + win[key].sum(0)
        m.add_samples(chosen)
        cand = cand[m.B:]
        if keep:
            u, c = torch.cat((batch, chosen)).unique(return_counts=True)
            cand = torch.cat((cand, u[c == 1]))
        picked += chosen.tolist()
    assert picked[:subset] == g["S"].tolist()                      # the replay itself reproduces the golden
    assert decided > 0


@pytest.mark.parametrize("name", CASES)
def test_free_running_selection_invariants(golden_dir, name):
    g = load(golden_dir, name)
    a = g["assignments"].astype(np.int64)
    keep = bool(g["keep_unselected"])
    V, subset = a.shape[0], int(g["subset"])
    gen_golden.seed_all(int(g["seed"]))
    m = gpu_measure(a, int(g["c"]), keep)
    m.init([tuple(p) for p in g["pairs"].tolist()], list(range(1, V)))
    S, GAIN, timelapse, LOOKUPS = m.run_greedy(subset, [0])
    assert len(S) == subset and len(set(S)) == subset and 0 not in S
    assert len(GAIN) >= subset and len(timelapse) == len(LOOKUPS)
    assert all(np.isfinite(GAIN)) and max(GAIN) < np.log(int(g["c"])) + 1e-3
    if keep:
        assert m.candidate_ids.shape[0] + len(GAIN) == V - 1
    # exact scoring (default): the same seed selects the very indices the unmodified reference selected, with its scores
    assert S == g["S"].tolist(), "free-running batch_mi must reproduce the reference's S"
    assert np.array_equal(np.array(GAIN, dtype=np.float32)[:len(g["GAIN"])], g["GAIN"].astype(np.float32)[:len(GAIN)])
