"""GPU parity: CUDA greedy-MI engine (through the C ABI) vs reference goldens and the C oracle."""
import os
import random

import numpy as np
import pytest
import torch

from acav100m_b200 import synth
from oracle import gen_golden, mi_oracle as mo

pytestmark = pytest.mark.gpu

P1 = [m for m in sorted(gen_golden.MI_CASES) if gen_golden.MI_CASES[m]["dcols"] == 2]
LOOPS = ["kernels", "persistent", "cells", "bytes"]


def load(golden_dir, name):
    return dict(np.load(os.path.join(golden_dir, name + ".npz")))


def gpu_measure(assignments, C, **kw):
    from acav100m_b200.subset_selection import get_measure
    return get_measure("mem_mi")(assignments, ncentroids=C, batch_size=20, selection_size=4, device="cuda",
                                 keep_unselected=True, **kw)


@pytest.mark.parametrize("loop", LOOPS)
@pytest.mark.parametrize("name", P1)
def test_selection_matches_reference_bits(golden_dir, name, loop):
    g = load(golden_dir, name + "_mem_mi")
    a = g["assignments"].astype(np.int64)
    order = g["candidate_order"].tolist()
    m = gpu_measure(a, int(g["c"]), loop=loop)
    m.init([tuple(p) for p in g["pairs"].tolist()], order[1:])
    S, GAIN, timelapse, LOOKUPS = m.run_greedy(int(g["subset"]), [order[0]])
    assert S == g["S"].tolist()
    assert np.array_equal(np.array(GAIN), g["GAIN"]), "fp32 scores must be bit-identical"
    assert len(timelapse) == len(GAIN) == len(LOOKUPS) == int(g["subset"]) - 2


@pytest.mark.parametrize("loop", LOOPS)
@pytest.mark.parametrize("W,C,picks,seed", [(20_000, 64, 3000, 1), (100_003, 256, 1500, 2), (5000, 1024, 800, 3),
                                            (257, 3, 256, 4)])
def test_selection_matches_c_oracle(W, C, picks, seed, loop):
    a = synth.zipf_pairs(W, C, seed)
    a[0] = C - 1
    pos_want, gain_want = mo.greedy_mem_mi_c(a[:, 0], a[:, 1], C, picks, bucketed=True)
    m = gpu_measure(a, C, loop=loop)
    m.init([(0, 1)], list(range(W)))
    pos, gain = m.select(picks)
    assert np.array_equal(pos.cpu().numpy(), pos_want)
    assert np.array_equal(gain.cpu().numpy(), gain_want)
    N, ca, rb, sums = m.read_state()
    assert int(N.sum()) == picks and int(ca.sum()) == picks and int(rb.sum()) == picks
    want_N = np.zeros((C, C), dtype=np.int64)
    np.add.at(want_N, (a[pos_want, 0], a[pos_want, 1]), 1)
    assert np.array_equal(N.numpy(), want_N)
    assert float(sums[3]) == float(picks)


def test_driver_matches_oracle_driver_with_shuffle():
    import types
    from acav100m_b200.subset_selection.run_greedy import _run_greedy
    a = synth.zipf_pairs(3000, 32, 77)
    a[5] = 31
    args = types.SimpleNamespace(batch=types.SimpleNamespace(batch_size=20, selection_size=4, keep_unselected=True),
                                 computation=types.SimpleNamespace(device="cuda"), log_every=1, log_times=None,
                                 node_rank=None, parent_pid=None)
    keys = [("audio", "layer_4"), ("video", "layer_4")]
    random.seed(5)
    S, GAIN, _ = _run_greedy(args, a, keys, None, 0.1, measure_name="mem_mi", shuffle_candidates=True)
    S2, GAIN2 = mo.run_greedy_driver(a, subset_ratio=0.1, shuffle_candidates=True, rng=random.Random(5))
    assert S == S2 and GAIN == GAIN2 and len(S) == 299


def test_add_samples_counts_into_table():
    a = synth.zipf_pairs(500, 8, 3)
    m = gpu_measure(a, 8)
    m.init([(0, 1)], list(range(10, 500)))
    m.add_samples(list(range(10)))
    N, ca, rb, sums = m.read_state()
    want = np.zeros((8, 8), dtype=np.int64)
    np.add.at(want, (a[:10, 0], a[:10, 1]), 1)
    assert np.array_equal(N.numpy(), want) and float(sums[3]) == 10.0


def test_two_engines_sharded_equal_one_engine():
    """The multi-GPU protocol (contiguous shards, pos_base, key max, broadcast apply) driven by hand
    on one device: two engines holding halves of the list pick exactly what one engine picks."""
    from acav100m_b200 import _lib
    from acav100m_b200.subset_selection.measures import tables
    W, C, picks = 7001, 16, 600
    a = torch.from_numpy(synth.zipf_pairs(W, C, 11))
    single = gpu_measure(a.numpy(), C)
    single.init([(0, 1)], list(range(W)))
    pos_want, gain_want = single.select(picks)
    logs = tables.log_table(W + 16).cuda()
    consts = tables.empty_table_constants(C)
    st = _lib.stream_ptr()
    engines, bounds = [], [(0, 3333), (3333, W)]
    cells = a.cuda()
    for lo, hi in bounds:
        h = _lib.c_vp()
        _lib.call("acav_mi_create", _lib.ctypes.byref(h), hi - lo, C, C, W + 8, lo)
        part = cells[lo:hi].contiguous()
        _lib.call("acav_mi_load_candidates", h, _lib.ptr(part), st)
        _lib.call("acav_mi_set_tables", h, _lib.ptr(logs), logs.numel(), consts.ctypes.data_as(_lib.c_vp), st)
        engines.append(h)
    pairs = torch.zeros(2, 2, dtype=torch.int64, device="cuda")
    pos = torch.empty(2, picks, dtype=torch.int64, device="cuda")
    gain = torch.empty(2, picks, dtype=torch.float32, device="cuda")
    for i in range(picks):
        for r, h in enumerate(engines):
            _lib.call("acav_mi_local_best", h, _lib.c_vp(pairs.data_ptr() + 16 * r), st)
        for r, h in enumerate(engines):
            _lib.call("acav_mi_apply", h, _lib.ptr(pairs), 2, _lib.c_vp(pos.data_ptr() + 8 * (r * picks + i)),
                      _lib.c_vp(gain.data_ptr() + 4 * (r * picks + i)), st)
    for h in engines:
        _lib.call("acav_mi_destroy", h)
    assert torch.equal(pos[0], pos_want) and torch.equal(pos[1], pos_want)
    assert torch.equal(gain[0], gain_want) and torch.equal(gain[1], gain_want)


def test_errors():
    a = synth.zipf_pairs(50, 4, 1)
    m = gpu_measure(a, 4)
    with pytest.raises(ValueError):
        m.init([(0, 1, 1)], list(range(50)))
    m.init([(0, 1)], list(range(1, 50)))
    with pytest.raises(RuntimeError):
        m.run_greedy(80, [0])
    bad = gpu_measure(a, 3)
    with pytest.raises(ValueError):
        bad.init([(0, 1)], list(range(50)))
    from acav100m_b200.subset_selection import get_measure
    with pytest.raises(AssertionError):
        get_measure("nope")


@pytest.mark.parametrize("W,C,picks,seed", [(50_000, 16, 4000, 21), (300_000, 1024, 600, 22), (40, 4, 39, 23),
                                            (1_000_003, 256, 300, 24), (9000, 2048, 500, 25)])
@pytest.mark.parametrize("loop", ["persistent", "cells", "bytes"])
def test_persistent_loop_matches_c_oracle(W, C, picks, seed, loop):
    """Row-partitioned persistent kernel: massive early ties (every cell scores the same at first),
    rows split across CTAs, more rows per CTA than fit in shared memory, resumed runs."""
    a = synth.zipf_pairs(W, C, seed)
    a[0] = C - 1
    pos_want, gain_want = mo.greedy_mem_mi_c(a[:, 0], a[:, 1], C, picks, bucketed=True)
    m = gpu_measure(a, C, loop=loop)
    m.init([(0, 1)], list(range(W)))
    first = picks // 3
    p1, g1 = m.select(first)
    p2, g2 = m.select(picks - first)                       # resumes from the engine's state
    pos = torch.cat([p1, p2]).cpu().numpy()
    gain = torch.cat([g1, g2]).cpu().numpy()
    assert np.array_equal(pos, pos_want)
    assert np.array_equal(gain, gain_want)


def test_loops_can_be_mixed():
    W, C, picks = 30_000, 32, 900
    a = synth.zipf_pairs(W, C, 31)
    pos_want, gain_want = mo.greedy_mem_mi_c(a[:, 0], a[:, 1], C, picks, bucketed=True)
    m = gpu_measure(a, C, loop="persistent")
    m.init([(0, 1)], list(range(W)))
    out = []
    for loop, n in (("persistent", 150), ("bytes", 100), ("cells", 100), ("kernels", 150), ("bytes", 100), ("cells", 100),
                    ("persistent", 200)):
        m.loop = loop
        out.append(m.select(n))
    pos = torch.cat([o[0] for o in out]).cpu().numpy()
    gain = torch.cat([o[1] for o in out]).cpu().numpy()
    assert np.array_equal(pos, pos_want) and np.array_equal(gain, gain_want)


@pytest.mark.parametrize("loop", ["persistent", "cells", "bytes"])
def test_persistent_loop_uniform_ids_all_ties(loop):
    """Every candidate in one cell: all scores tie on every iteration, list order must be kept."""
    W = 5000
    a = np.zeros((W, 2), dtype=np.int64)
    a[:, 1] = 3
    a[0] = (7, 7)
    m = gpu_measure(a, 8, loop=loop)
    m.init([(0, 1)], list(range(W)))
    pos, _ = m.select(200)
    want, _ = mo.greedy_mem_mi_c(a[:, 0], a[:, 1], 8, 200, bucketed=False)
    assert np.array_equal(pos.cpu().numpy(), want)


@pytest.mark.parametrize("C,want_loop", [(8140, "persistent"), (8192, "cells"), (16384, "cells")])
def test_auto_loop_falls_back_when_the_table_outgrows_shared_memory(C, want_loop):
    """loop='auto' (the CLI default) must pick a loop that can run the table: the persistent stream needs one gain
    row of K_v + 1 floats next to the replicated marginals in shared memory (K <= 8140), the cell index needs the
    marginals (K <= 16384) -- and the picks stay the oracle's.  (Beyond K = 16384 the empty table's n0 = K^2 * eps is
    no longer absorbed by fp32 1.0 and the log-table trick does not apply: tables.py refuses such tables.)"""
    W, picks = 6000, 12
    a = synth.zipf_pairs(W, C, C)
    a[0] = C - 1
    pos_want, gain_want = mo.greedy_mem_mi_c(a[:, 0], a[:, 1], C, picks, bucketed=True)
    m = gpu_measure(a, C)                                    # loop='auto'
    m.init([(0, 1)], list(range(W)))
    assert m.loop_name() == want_loop
    pos, gain = m.select(picks)
    assert np.array_equal(pos.cpu().numpy(), pos_want)
    assert np.array_equal(gain.cpu().numpy(), gain_want)
    if want_loop != "persistent":                            # asking for a loop that cannot run the shape is an error
        from acav100m_b200 import _lib
        bad = gpu_measure(a, C, loop="persistent")
        bad.init([(0, 1)], list(range(W)))
        with pytest.raises(_lib.AcavError) as e:
            bad.select(1)
        assert e.value.status == -2


@pytest.mark.parametrize("variant,cache", [(0, 1), (1, 1), (2, 1), (3, 1), (4, 1), (5, 1)])
@pytest.mark.parametrize("W,C,picks,seed", [(120_000, 300, 700, 5), (60_000, 1024, 400, 6)])
def test_byte_stream_variants_match_c_oracle(W, C, picks, seed, variant, cache):
    """Every (threads, loads in flight) instantiation of the one-byte stream kernel: same picks, same fp32 gains.
    C = 300 -> 2 sub-rows of 150 columns, C = 1024 -> 5 of 205 (5120 sub-rows, most of them a single padded block:
    the interleaved sub-row order and the parked-tie list are both exercised)."""
    from acav100m_b200 import _lib
    a = synth.zipf_pairs(W, C, seed)
    a[0] = C - 1
    pos_want, gain_want = mo.greedy_mem_mi_c(a[:, 0], a[:, 1], C, picks, bucketed=True)
    m = gpu_measure(a, C, loop="bytes")
    m.init([(0, 1)], list(range(W)))
    _lib.call("acav_mi_set_stream_variant", m._engine, variant, cache)
    p1, g1 = m.select(picks // 3)                             # two launches: state written back and re-read
    p2, g2 = m.select(picks - picks // 3)
    pos, gain = torch.cat([p1, p2]).cpu().numpy(), torch.cat([g1, g2]).cpu().numpy()
    assert np.array_equal(pos, pos_want)
    assert np.array_equal(gain, gain_want)
    N, ca, rb, sums = m.read_state()
    want_N = np.zeros((C, C), dtype=np.int64)
    np.add.at(want_N, (a[pos_want, 0], a[pos_want, 1]), 1)
    assert np.array_equal(N.numpy(), want_N)


@pytest.mark.parametrize("W,C,picks,want_loop", [(30_000, 64, 300, "cells"), (60_000, 4096, 60, None)])
def test_auto_loop_picks_a_supported_loop_and_matches_c_oracle(W, C, picks, want_loop):
    """loop='auto' (the CLI default) orders the exact loops by their measured cost for the shape and takes the first
    one whose layout can be built for this list: the cell index wherever K^2 << W; at K = 4096 neither the one-byte
    stream (69632 sub-rows) nor the cell index fits the cost model's first place, whatever runs must give the oracle's picks."""
    from acav100m_b200 import _lib
    a = synth.zipf_pairs(W, C, 41)
    pos_want, gain_want = mo.greedy_mem_mi_c(a[:, 0], a[:, 1], C, picks, bucketed=True)
    m = gpu_measure(a, C, loop="auto")
    m.init([(0, 1)], list(range(W)))
    order = m._auto_order()
    assert order[-1] == _lib.MI_LOOP_KERNELS and all(_lib.load().acav_mi_loop_supported(C, C, x) for x in order)
    pos, gain = m.select(picks)
    assert np.array_equal(pos.cpu().numpy(), pos_want) and np.array_equal(gain.cpu().numpy(), gain_want)
    if want_loop is not None:
        assert m.loop_name() == want_loop


def test_auto_loop_moves_on_when_a_layout_cannot_be_built(monkeypatch):
    """acav_mi_prepare returning ACAV_E_UNSUPPORTED for the first choice (here: forced) makes select() take the next loop."""
    from acav100m_b200 import _lib
    W, C, picks = 20_000, 32, 120
    a = synth.zipf_pairs(W, C, 43)
    pos_want, gain_want = mo.greedy_mem_mi_c(a[:, 0], a[:, 1], C, picks, bucketed=True)
    m = gpu_measure(a, C, loop="auto")
    m.init([(0, 1)], list(range(W)))
    first = m._auto_order()[0]
    real_call = _lib.call

    def flaky(name, *args):
        if name == "acav_mi_prepare" and args[1] == first:
            raise _lib.AcavError(name, _lib.E_UNSUPPORTED, "forced by the test")
        return real_call(name, *args)

    monkeypatch.setattr(_lib, "call", flaky)
    pos, gain = m.select(picks)
    assert m._loop_mode() == m._auto_order()[1]
    assert np.array_equal(pos.cpu().numpy(), pos_want) and np.array_equal(gain.cpu().numpy(), gain_want)
