"""Pin the oracle (oracle/*.py, oracle/mi_oracle.c) against outputs of the reference itself.

The fixtures under tests/golden/ were written by oracle/gen_golden.py, which runs the UNMODIFIED
reference (sgd_clustering.py KMeans; measures/mi.py EfficientMemMI / EfficientMI) seeded on CPU.
Bit-exactness is asserted wherever the oracle calls the same torch CPU operators as the reference.
When /root/reference is mounted the restatements are additionally re-checked live.
"""
import glob
import os
import random

import numpy as np
import pytest
import torch

from acav100m_b200 import synth
from oracle import gen_golden, kmeans_oracle as ko, mi_oracle as mo, ref_shims

KM = sorted(gen_golden.KMEANS_CASES)
MI = sorted(gen_golden.MI_CASES)


def load(golden_dir, name):
    return dict(np.load(os.path.join(golden_dir, name + ".npz")))


def test_all_fixtures_present(golden_dir):
    names = {os.path.basename(p)[:-4] for p in glob.glob(os.path.join(golden_dir, "*.npz"))}
    want = set(KM) | {m + "_mem_mi" for m in MI} | {"mi_dense_small_mi"}
    assert want <= names


@pytest.mark.parametrize("name", KM)
def test_kmeans_oracle_reproduces_reference_bits(golden_dir, name):
    case, g = gen_golden.KMEANS_CASES[name], load(golden_dir, name)
    x = torch.from_numpy(synth.gaussian_mixture(case["n"], case["d"], case["k_true"], case["seed"]))
    assert float(x.double().sum()) == float(g["x_checksum"]), "synthetic input stream drifted"
    gen_golden.seed_all(case["seed"])
    st = ko.new_state(case["d"], case["k"])
    st.sequential = bool(case.get("sequential", False))
    assert np.array_equal(st.centers.numpy(), g["init_centers"])
    batches = gen_golden.kmeans_batches(x, case["batch"])
    dists = ko.train(st, lambda epoch: batches, case["epochs"])
    assert np.array_equal(st.centers.numpy(), g["centers"])          # bit-exact
    assert np.array_equal(st.counts.numpy(), g["counts"])
    assert st.count == int(g["count"]) and st.fallback == int(g["fallback"])
    assert np.array_equal(np.array(dists), g["step_mean_dist"])
    best, mean_d = ko.assign(st, x)
    assert np.array_equal(best.numpy(), g["assign_best"])
    assert mean_d == float(g["assign_mean_dist"])


@pytest.mark.parametrize("name", KM)
def test_kmeans_fp64_truth_agrees_outside_ambiguity_band(golden_dir, name):
    case, g = gen_golden.KMEANS_CASES[name], load(golden_dir, name)
    x = synth.gaussian_mixture(case["n"], case["d"], case["k_true"], case["seed"])
    st = ko.SgdKMeansState(centers=torch.from_numpy(g["centers"]), counts=torch.from_numpy(g["counts"]),
                           count=int(g["count"]))
    mask = ko.underused_mask(st).numpy()
    best64, d1, d2 = ko.assign_truth_f64(g["centers"], x, mask)
    differ = best64 != g["assign_best"]
    band = ko.fp32_ambiguity_band(g["centers"], x)
    assert np.all((d2 - d1)[differ] <= band[differ])
    assert differ.mean() < 0.01


def test_scatter_add_is_strict_row_order():
    """The CUDA segmented sum reproduces row order; check torch's CPU scatter_add_ (the oracle's and
    the shimmed reference's definition) is that same order (torch-scatter 2.0.5 semantics)."""
    rng = np.random.RandomState(5)
    src = (rng.standard_normal((3000, 7)) * 10 ** rng.uniform(-3, 3, (3000, 1))).astype(np.float32)
    idx = rng.randint(0, 11, size=3000)
    want = ko.sequential_scatter_sum(src, idx, 11)
    got = torch.zeros(11, 7).scatter_add_(0, torch.from_numpy(idx)[:, None].expand(-1, 7),
                                          torch.from_numpy(src)).numpy()
    assert np.array_equal(want, got)


def _mi_case_inputs(g):
    a = g["assignments"].astype(np.int64)
    order = g["candidate_order"]
    return a, [int(order[0])], [int(i) for i in order[1:]], [tuple(p) for p in g["pairs"].tolist()]


@pytest.mark.parametrize("name", MI)
def test_mem_mi_torch_restatement_reproduces_reference_bits(golden_dir, name):
    g = load(golden_dir, name + "_mem_mi")
    a, start, cands, pairs = _mi_case_inputs(g)
    S, GAIN = mo.greedy_mem_mi(a, int(g["c"]), pairs, cands, int(g["subset"]), start)
    assert S == g["S"].tolist()
    assert np.array_equal(np.array(GAIN), g["GAIN"])                 # bit-exact fp32 scores
    assert len(S) == int(g["subset"]) - 1                            # mi.py:161 off-by-one kept


@pytest.mark.parametrize("bucketed", [False, True])
@pytest.mark.parametrize("name", [m for m in MI if gen_golden.MI_CASES[m]["dcols"] == 2])
def test_mem_mi_c_restatement_reproduces_reference_bits(golden_dir, name, bucketed):
    g = load(golden_dir, name + "_mem_mi")
    a, start, cands, pairs = _mi_case_inputs(g)
    assert pairs == [(0, 1)]
    cells = a[np.array(cands)]
    pos, gain = mo.greedy_mem_mi_c(cells[:, 0], cells[:, 1], int(g["c"]), int(g["subset"]) - 2,
                                   bucketed=bucketed)
    S = start + [cands[p] for p in pos]
    assert S == g["S"].tolist()
    assert np.array_equal(gain.astype(np.float64), g["GAIN"])


def test_mem_mi_driver_matches_golden(golden_dir):
    g = load(golden_dir, "mi_c1_mem_mi")
    S, GAIN = mo.run_greedy_driver(g["assignments"].astype(np.int64), subset_size=int(g["subset"]))
    assert S == g["S"].tolist()
    S2, GAIN2 = mo.greedy_mem_mi_via_c(g["assignments"].astype(np.int64), int(g["subset"]))
    assert S2 == S and np.array_equal(np.array(GAIN2), np.array(GAIN))


def test_dense_mi_and_mem_mi_goldens_differ_only_by_tie_breaking(golden_dir):
    """SURVEY headline fact 3: `mi` and `mem_mi` do not select the same sequence (fp noise vs index
    tie-breaks); the fixtures document it, and both sets have the reference's |S| = subset-1."""
    a, b = load(golden_dir, "mi_dense_small_mi"), load(golden_dir, "mi_dense_small_mem_mi")
    assert len(a["S"]) == len(b["S"]) == int(a["subset"]) - 1
    assert a["S"].tolist() != b["S"].tolist()


def test_torch_log_is_position_independent():
    """log_table() stands in for x.log() evaluated inside [W, P] tensors: same bits anywhere."""
    tab = mo.log_table(200_000)
    perm = torch.randperm(200_000, generator=torch.Generator().manual_seed(0)) + 1
    got = perm.to(torch.float32).log().numpy()
    assert np.array_equal(got, tab[perm.numpy()])
    assert torch.tensor(77.0).log().item() == float(tab[77])
    odd = torch.arange(1, 38, dtype=torch.float32)[::3].log().numpy()
    assert np.array_equal(odd, tab[1:38:3])


def test_pairing_matches_reference_definitions():
    keys = [("a", "l0"), ("a", "l1"), ("v", "l0"), ("v", "l1")]
    assert mo.cluster_pairing(keys, "combination") == [(0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3)]
    assert mo.cluster_pairing(keys, "bipartite") == [(0, 2), (0, 3), (1, 2), (1, 3)]
    assert mo.cluster_pairing(keys, "diagonal") == [[0, 2], [1, 3]]


@pytest.mark.skipif(not ref_shims.reference_available(), reason="/root/reference not mounted")
def test_live_reference_agrees_with_oracle_on_fresh_seeds():
    """Beyond the committed fixtures: fresh seeds, reference run live (container only)."""
    get_measure, get_pairing = ref_shims.load_reference_measures()
    for seed in (11, 12):
        a = synth.zipf_pairs(500, 12, seed)
        a[0] = 11
        gen_golden.seed_all(seed)
        m = get_measure("mem_mi")(a, ncentroids=12, batch_size=20, selection_size=4, device="cpu",
                                  keep_unselected=True)
        cands = list(range(500))
        random.Random(seed).shuffle(cands)
        m.init([(0, 1)], cands[1:])
        S, GAIN, _, _ = m.run_greedy(120, [cands[0]])
        S2, GAIN2 = mo.greedy_mem_mi(a, 12, [(0, 1)], cands[1:], 120, [cands[0]])
        assert S == S2 and GAIN == GAIN2
    KMeans = ref_shims.load_reference_kmeans()
    x = torch.from_numpy(synth.gaussian_mixture(640, 48, 6, 21))
    batches = gen_golden.kmeans_batches(x, 64)
    gen_golden.seed_all(21)                     # both draw init + warm-up noise from the global RNG
    km = KMeans(ref_shims.reference_kmeans_args(), 48, 8)
    with ref_shims.cuda_is_identity():
        d_ref = [km.add(xb) for xb in batches]
    gen_golden.seed_all(21)
    st = ko.new_state(48, 8)
    d_or = [ko.sgd_step(st, xb)[1] for xb in batches]
    assert d_ref == d_or
    assert torch.equal(km.centers, st.centers) and torch.equal(km.counts, st.counts)


def test_dense_mi_restatement_reproduces_reference_run(golden_dir):
    """oracle.batch_mi_oracle.greedy_dense_mi (mi.py:150-192 with the dense calc_MI) against the reference's own
    `mi` run: same picks, same fp32 scores."""
    from oracle import batch_mi_oracle as bo
    g = load(golden_dir, "mi_dense_small_mi")
    a = g["assignments"].astype(np.int64)
    order = g["candidate_order"].tolist()
    S, GAIN, _ = bo.greedy_dense_mi(a, int(g["c"]), [tuple(p) for p in g["pairs"].tolist()], order[1:],
                                    int(g["subset"]), [order[0]])
    assert S == g["S"].tolist()
    assert np.array_equal(np.array(GAIN), g["GAIN"])
