"""GPU parity of the multi-pair (P > 1) greedy-MI engine (acav_mi_pairs_*, csrc/mi_pairs.cu) through the reference's
measure API: goldens written by the unmodified reference (P = 3, 10, 45) and the C oracle at larger sizes.  Picks and
fp32 scores must be bit-identical (the mean over pairs is added in torch's CPU order)."""
import os

import numpy as np
import pytest
import torch

from oracle import gen_golden, mi_oracle as mo

pytestmark = pytest.mark.gpu

PAIR_CASES = sorted(n for n, c in gen_golden.MI_CASES.items() if c["dcols"] > 2)


def gpu_measure(assignments, C, **kw):
    from acav100m_b200.subset_selection import get_measure
    return get_measure("mem_mi")(assignments, ncentroids=C, batch_size=20, selection_size=4, device="cuda",
                                 keep_unselected=True, **kw)


def combos(d, limit=None):
    pairs = mo.cluster_pairing([("m%d" % i, "l") for i in range(d)], "combination")
    return pairs[:limit] if limit else pairs


def oracle_after_forced_samples(a, C, pairs, n_forced, picks):
    """torch restatement: count the first `n_forced` rows into the tables (add_samples, mi.py:408-412), then pick
    greedily among the rest; positions are relative to the first remaining row."""
    tab = mo.init_table(len(pairs), C)
    cells = mo.candidate_cells(a, pairs, list(range(len(a))))
    for i in range(n_forced):
        _, NlogN, aloga, blogb = mo.candidate_scores(tab, cells[i:i + 1])
        mo.apply_pick(tab, cells[i], NlogN[0], aloga[0], blogb[0])
    ids, rest = torch.arange(n_forced, len(a)), cells[n_forced:]
    want_pos, want_gain = [], []
    for _ in range(picks):
        scores, NlogN, aloga, blogb = mo.candidate_scores(tab, rest)
        score, idx = scores.mean(dim=-1).max(dim=0)
        idx = idx.item()
        want_pos.append(int(ids[idx]) - n_forced)
        want_gain.append(score.item())
        mo.apply_pick(tab, rest[idx], NlogN[idx], aloga[idx], blogb[idx])
        keep = torch.ones(len(ids), dtype=torch.bool)
        keep[idx] = False
        ids, rest = ids[keep], rest[keep]
    return want_pos, want_gain


@pytest.mark.parametrize("name", PAIR_CASES)
def test_selection_matches_reference_bits(golden_dir, name):
    g = dict(np.load(os.path.join(golden_dir, name + "_mem_mi.npz")))
    a = g["assignments"].astype(np.int64)
    order = g["candidate_order"].tolist()
    m = gpu_measure(a, int(g["c"]))
    m.init([tuple(p) for p in g["pairs"].tolist()], order[1:])
    assert m.loop_name() == "pairs"
    S, GAIN, timelapse, LOOKUPS = m.run_greedy(int(g["subset"]), [order[0]])
    assert S == g["S"].tolist()
    assert np.array_equal(np.array(GAIN), g["GAIN"]), "fp32 scores must be bit-identical"
    assert len(timelapse) == len(GAIN) == len(LOOKUPS) == int(g["subset"]) - 2


@pytest.mark.parametrize("W,D,C,npairs,picks,seed", [
    (20_000, 10, 16, None, 120, 1),        # the reference default: ten clusterings, 45 pairs
    (100_003, 4, 64, None, 200, 2),        # 6 pairs (scalar summation path), W not a multiple of the block
    (5_000, 23, 7, 250, 80, 3),            # 250 pairs: 31 vectors + 2 trailing values in the mean
    (3_000, 6, 300, 9, 60, 4),             # larger tables, 9 pairs (one vector + one trailing value)
    (41, 3, 4, None, 40, 5),               # the list runs empty
])
def test_selection_matches_c_oracle(W, D, C, npairs, picks, seed):
    rng = np.random.RandomState(seed)
    base = rng.randint(0, C, size=(W, 1))
    a = np.where(rng.random_sample((W, D)) < 0.5, (base + np.arange(D)) % C, rng.randint(0, C, size=(W, D))).astype(np.int64)
    pairs = combos(D, npairs)
    pos_want, gain_want, sums_want = mo.greedy_mem_mi_pairs_c(a, C, pairs, picks, return_sums=True)
    m = gpu_measure(a, C)
    m.init(pairs, list(range(W)))
    first = picks // 3
    p1, g1 = m.select(first)
    p2, g2 = m.select(picks - first)                       # resumes from the engine's state
    pos = torch.cat([p1, p2]).cpu().numpy()
    gain = torch.cat([g1, g2]).cpu().numpy()
    assert np.array_equal(pos, pos_want)
    assert np.array_equal(gain, gain_want)
    N, ca, rb, sums = m.read_state()
    assert np.array_equal(sums.numpy(), sums_want)
    for p, (c1, c2) in enumerate(pairs):
        want_N = np.zeros((C, C), dtype=np.int64)
        np.add.at(want_N, (a[pos_want, c1], a[pos_want, c2]), 1)
        assert np.array_equal(N[p].numpy(), want_N)
        assert np.array_equal(ca[p].numpy(), want_N.sum(0)) and np.array_equal(rb[p].numpy(), want_N.sum(1))


def test_pairs_engine_with_one_pair_equals_the_p1_engine():
    from acav100m_b200.subset_selection.measures.pairs_engine import PairsEngine
    rng = np.random.RandomState(6)
    a = rng.randint(0, 32, size=(30_000, 2)).astype(np.int64)
    m = gpu_measure(a, 32)
    m.init([(0, 1)], list(range(len(a))))
    pos_want, gain_want = m.select(500)
    eng = PairsEngine(torch.device("cuda", torch.cuda.current_device()), 32, [(0, 1)], torch.from_numpy(a), 0, len(a),
                      len(a) + 8)
    pos, gain = eng.select(500)
    eng.release()
    assert torch.equal(pos, pos_want) and torch.equal(gain, gain_want)


def test_subset_of_columns_and_add_samples():
    """Pairs that mention only some clustering columns (bipartite / diagonal pairings), samples counted in first."""
    rng = np.random.RandomState(7)
    a = rng.randint(0, 9, size=(4000, 6)).astype(np.int64)
    pairs = [(1, 4), (5, 1), (4, 5)]                       # columns 0, 2, 3 unused; (5, 1) is not sorted
    m = gpu_measure(a, 9)
    m.init(pairs, list(range(10, 4000)))
    m.add_samples(list(range(10)))
    pos, gain = m.select(100)
    want_pos, want_gain = oracle_after_forced_samples(a, 9, pairs, 10, 100)
    assert pos.cpu().tolist() == want_pos
    assert np.array_equal(gain.cpu().numpy(), np.array(want_gain, dtype=np.float32))


def test_two_engines_sharded_equal_one_engine():
    """The multi-GPU protocol of the pairs engine (contiguous shards, records = key + ids, every engine applies the
    best record) driven by hand on one device."""
    from acav100m_b200 import _lib
    from acav100m_b200.subset_selection.measures import tables
    rng = np.random.RandomState(8)
    W, D, C, picks = 6001, 5, 12, 150
    a = rng.randint(0, C, size=(W, D)).astype(np.int64)
    pairs = combos(D)
    pos_want, gain_want = mo.greedy_mem_mi_pairs_c(a, C, pairs, picks)
    pr = np.ascontiguousarray(pairs, dtype=np.int32)
    logs = tables.log_table(W + 16).cuda()
    consts = np.ascontiguousarray(tables.pair_table_constants(len(pairs), C))
    st = _lib.stream_ptr()
    rows = torch.from_numpy(a).cuda()
    engines = []
    for lo, hi in [(0, 2500), (2500, W)]:
        h = _lib.c_vp()
        _lib.call("acav_mi_pairs_create", _lib.ctypes.byref(h), hi - lo, D, C, len(pairs), pr.ctypes.data_as(_lib.c_vp),
                  W + 8, lo)
        part = rows[lo:hi].contiguous()
        _lib.call("acav_mi_pairs_load_candidates", h, _lib.ptr(part), st)
        _lib.call("acav_mi_pairs_set_tables", h, _lib.ptr(logs), logs.numel(), consts.ctypes.data_as(_lib.c_vp), st)
        engines.append(h)
    words = _lib.load().acav_mi_pairs_record_words(engines[0])
    assert words == 1 + (D + 3) // 4
    recs = torch.zeros(2, words, dtype=torch.int64, device="cuda")
    pos = torch.empty(2, picks, dtype=torch.int64, device="cuda")
    gain = torch.empty(2, picks, dtype=torch.float32, device="cuda")
    for i in range(picks):
        for r, h in enumerate(engines):
            _lib.call("acav_mi_pairs_local_best", h, _lib.c_vp(recs.data_ptr() + 8 * words * r), st)
        for r, h in enumerate(engines):
            _lib.call("acav_mi_pairs_apply", h, _lib.ptr(recs), 2, _lib.c_vp(pos.data_ptr() + 8 * (r * picks + i)),
                      _lib.c_vp(gain.data_ptr() + 4 * (r * picks + i)), st)
    torch.cuda.synchronize()
    for h in engines:
        _lib.call("acav_mi_pairs_destroy", h)
    for r in range(2):
        assert np.array_equal(pos[r].cpu().numpy(), pos_want)
        assert np.array_equal(gain[r].cpu().numpy(), gain_want)


def test_driver_with_default_pairing_matches_oracle_driver():
    """`_run_greedy` as the CLI calls it: ten clustering columns, `combination` pairing (P = 45), no shuffle."""
    import types
    from acav100m_b200.subset_selection.run_greedy import _run_greedy
    rng = np.random.RandomState(9)
    a = rng.randint(0, 6, size=(600, 10)).astype(np.int64)
    a[3] = 5
    args = types.SimpleNamespace(batch=types.SimpleNamespace(batch_size=20, selection_size=4, keep_unselected=True),
                                 computation=types.SimpleNamespace(device="cuda"), log_every=1, log_times=None,
                                 node_rank=None, parent_pid=None)
    keys = [("m%d" % i, "layer") for i in range(10)]
    S, GAIN, _ = _run_greedy(args, a, keys, 40, None, measure_name="mem_mi", shuffle_candidates=False)
    S2, GAIN2 = mo.run_greedy_driver(a, subset_size=40, clustering_types=keys)
    assert S == S2 and GAIN == GAIN2


def test_limits_are_reported():
    from acav100m_b200 import _lib
    h = _lib.c_vp()
    pr = np.zeros((1, 2), dtype=np.int32)
    with pytest.raises(_lib.AcavError) as e:
        _lib.call("acav_mi_pairs_create", _lib.ctypes.byref(h), 10, 65, 4, 1, pr.ctypes.data_as(_lib.c_vp), 8, 0)
    assert e.value.status == -2
    bad = np.array([[0, 7]], dtype=np.int32)
    with pytest.raises(_lib.AcavError) as e:
        _lib.call("acav_mi_pairs_create", _lib.ctypes.byref(h), 10, 3, 4, 1, bad.ctypes.data_as(_lib.c_vp), 8, 0)
    assert e.value.status == -1


def test_init_from_cells_with_several_pairs_equals_init():
    """The big-list setup (no list(range(V)), no per-candidate python objects) for P > 1: same picks and scores as
    init(pairs, candidates), and the unshuffled driver takes it."""
    import types
    from acav100m_b200.subset_selection import get_measure
    from acav100m_b200.subset_selection.run_greedy import _run_greedy
    rng = np.random.RandomState(11)
    V, D, C, picks = 20_001, 5, 16, 120
    a = rng.randint(0, C, size=(V, D)).astype(np.int64)
    pairs = mo.cluster_pairing([("m%d" % i, "l") for i in range(D)], "combination")          # P = 10
    want_pos, want_gain = mo.greedy_mem_mi_pairs_c(a[1:], C, pairs, picks)
    m = get_measure("mem_mi")(a, ncentroids=C, device="cuda")
    m.init_from_cells(pairs, torch.from_numpy(a[1:]), w_global=V - 1, id_offset=1)
    pos, gain = m.select(picks)
    assert np.array_equal(pos.cpu().numpy(), want_pos) and np.array_equal(gain.cpu().numpy(), want_gain)
    args = types.SimpleNamespace(batch=types.SimpleNamespace(batch_size=20, selection_size=4, keep_unselected=True),
                                 computation=types.SimpleNamespace(device="cuda"), log_every=1, log_times=None,
                                 node_rank=None, parent_pid=None)
    keys = [("m%d" % i, "l") for i in range(D)]
    S, GAIN, _ = _run_greedy(args, a, keys, picks + 2, None, measure_name="mem_mi", shuffle_candidates=False)
    assert S == [0] + (want_pos + 1).tolist() and np.array_equal(np.array(GAIN, dtype=np.float32), want_gain)
