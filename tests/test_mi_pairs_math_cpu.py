"""CPU checks of the multi-pair (P > 1) greedy-MI path that need no GPU.

* the C oracle's restatement of torch's inner-dimension sum order against torch itself;
* the C oracle's multi-pair greedy against the goldens written by the unmodified reference (P = 3, 10, 45);
* the arithmetic header the CUDA kernels are built from (csrc/mi_pairs_math.h), compiled for the host by
  tests/native/mi_pairs_math_host.cpp and walked through the kernels' data flow, against the oracle bit for bit.
"""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest
import torch

from oracle import gen_golden, mi_oracle as mo

HERE = os.path.dirname(os.path.abspath(__file__))
PAIR_CASES = sorted(n for n, c in gen_golden.MI_CASES.items() if c["dcols"] > 2)
f32p, i32p, i64p = (ctypes.POINTER(t) for t in (ctypes.c_float, ctypes.c_int32, ctypes.c_int64))


def _golden(golden_dir, name):
    g = dict(np.load(os.path.join(golden_dir, name + "_mem_mi.npz")))
    a = g["assignments"].astype(np.int64)
    order = g["candidate_order"]
    return g, a, int(order[0]), order[1:], [tuple(p) for p in g["pairs"].tolist()]


def _rows(rng, n, p):
    return (rng.standard_normal((n, p)) * rng.choice([1e-3, 1.0, 1e3], size=(n, p))).astype(np.float32)


def test_aten_row_mean_restatement_matches_torch():
    """mi_oracle.c restates ATen's cascade_sum for 8-lane vectors; torch's `mean(dim=-1)` must agree bit for bit for
    every P the engine admits and beyond (the cascade levels engage from P = 512)."""
    rng = np.random.RandomState(3)
    for p in list(range(1, 131)) + [255, 256, 257, 511, 512, 513, 600, 1024, 2100]:
        x = _rows(rng, 16, p)
        want = torch.from_numpy(x).mean(dim=-1).numpy()
        got = np.array([mo.aten_row_mean(r) for r in x], dtype=np.float32)
        assert np.array_equal(want, got), p


@pytest.mark.parametrize("threads", [1, 4])
def test_aten_row_mean_restatement_at_scale_and_any_thread_count(threads):
    """torch parallelises the reduction over output rows, never inside a row: the order holds for W = 3*10^5 rows and does
    not depend on the number of threads (the GPU box has twice the cores of the container the goldens were made in)."""
    rng = np.random.RandomState(5)
    before = torch.get_num_threads()
    torch.set_num_threads(threads)
    try:
        for w, p in [(300_000, 45), (400_000, 6)]:
            x = (rng.standard_normal((w, p)) * 5).astype(np.float32)
            want = torch.from_numpy(x).mean(dim=-1).numpy()
            idx = rng.randint(0, w, size=1500)
            got = np.array([mo.aten_row_mean(x[i]) for i in idx], dtype=np.float32)
            assert np.array_equal(want[idx], got), (w, p)
    finally:
        torch.set_num_threads(before)


@pytest.mark.parametrize("name", PAIR_CASES)
def test_c_pairs_oracle_reproduces_reference_bits(golden_dir, name):
    g, a, start, cands, pairs = _golden(golden_dir, name)
    pos, gain = mo.greedy_mem_mi_pairs_c(a[cands], int(g["c"]), pairs, int(g["subset"]) - 2)
    assert [start] + cands[pos].tolist() == g["S"].tolist()
    assert np.array_equal(gain.astype(np.float64), g["GAIN"])


def test_c_pairs_oracle_equals_torch_restatement_down_to_the_last_candidate():
    rng = np.random.RandomState(9)
    a = rng.randint(0, 5, size=(70, 4))
    a[0] = 4
    pairs = mo.cluster_pairing([("m%d" % i, "l") for i in range(4)], "combination")
    S, GAIN = mo.greedy_mem_mi(a, 5, pairs, list(range(1, 70)), 71, [0])          # 69 picks: the list runs empty
    pos, gain = mo.greedy_mem_mi_pairs_c(a[1:], 5, pairs, 69)
    assert S[1:] == (pos + 1).tolist()
    assert np.array_equal(np.array(GAIN, dtype=np.float32), gain)


def test_c_pairs_oracle_with_one_pair_equals_the_p1_oracle():
    rng = np.random.RandomState(10)
    a = rng.randint(0, 9, size=(500, 2))
    pos1, gain1 = mo.greedy_mem_mi_c(a[:, 0], a[:, 1], 9, 120, bucketed=False)
    pos, gain = mo.greedy_mem_mi_pairs_c(a, 9, [(0, 1)], 120)
    assert np.array_equal(pos, pos1) and np.array_equal(gain, gain1)


@pytest.fixture(scope="module")
def host_math(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    out = str(tmp_path_factory.mktemp("native") / "libmi_pairs_math_host.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared",
                           "-o", out, os.path.join(HERE, "native", "mi_pairs_math_host.cpp")])
    lib = ctypes.CDLL(out)
    lib.host_pairs_mean.restype = ctypes.c_float
    lib.host_pairs_mean.argtypes = [f32p, ctypes.c_int32]
    lib.host_pairs_greedy.restype = ctypes.c_int64
    lib.host_pairs_greedy.argtypes = [i32p, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, i32p, ctypes.c_int32, f32p,
                                      f32p, ctypes.c_int64, i64p, f32p, f32p]
    return lib


def test_device_header_mean_order_matches_torch(host_math):
    rng = np.random.RandomState(4)
    for p in range(1, 257):
        x = _rows(rng, 8, p)
        want = torch.from_numpy(x).mean(dim=-1).numpy()
        got = np.array([host_math.host_pairs_mean(r.ctypes.data_as(f32p), p) for r in np.ascontiguousarray(x)],
                       dtype=np.float32)
        assert np.array_equal(want, got), p


def _host_greedy(lib, ids, c, pairs, n_picks):
    ids = np.ascontiguousarray(ids, dtype=np.int32)
    pr = np.ascontiguousarray(np.asarray(pairs, dtype=np.int32).reshape(-1, 2))
    logs = mo.log_table(n_picks + 2)
    consts = np.ascontiguousarray(mo.pair_constants(len(pr), c))
    pos, gain = np.zeros(n_picks, dtype=np.int64), np.zeros(n_picks, dtype=np.float32)
    sums = np.zeros((len(pr), 4), dtype=np.float32)
    done = lib.host_pairs_greedy(ids.ctypes.data_as(i32p), ids.shape[0], ids.shape[1], c, pr.ctypes.data_as(i32p),
                                 len(pr), logs.ctypes.data_as(f32p), consts.ctypes.data_as(f32p), n_picks,
                                 pos.ctypes.data_as(i64p), gain.ctypes.data_as(f32p), sums.ctypes.data_as(f32p))
    assert done == n_picks
    return pos, gain, sums


@pytest.mark.parametrize("name", PAIR_CASES)
def test_device_header_data_flow_reproduces_reference_bits(host_math, golden_dir, name):
    g, a, start, cands, pairs = _golden(golden_dir, name)
    pos, gain, _ = _host_greedy(host_math, a[cands], int(g["c"]), pairs, int(g["subset"]) - 2)
    assert [start] + cands[pos].tolist() == g["S"].tolist()
    assert np.array_equal(gain.astype(np.float64), g["GAIN"])


@pytest.mark.parametrize("d,c,w,picks,seed", [(2, 7, 300, 100, 1), (6, 12, 2000, 150, 2), (10, 16, 1500, 60, 3),
                                             (23, 4, 400, 399, 4)])
def test_device_header_data_flow_equals_oracle(host_math, d, c, w, picks, seed):
    rng = np.random.RandomState(seed)
    ids = rng.randint(0, c, size=(w, d))
    pairs = mo.cluster_pairing([("m%d" % i, "l") for i in range(d)], "combination")[:256]
    want_pos, want_gain, want_sums = mo.greedy_mem_mi_pairs_c(ids, c, pairs, picks, return_sums=True)
    pos, gain, sums = _host_greedy(host_math, ids, c, pairs, picks)
    assert np.array_equal(pos, want_pos)
    assert np.array_equal(gain, want_gain)
    assert np.array_equal(sums, want_sums)


def test_device_header_data_flow_equals_oracle_on_random_instances(host_math):
    """Random shapes (2..23 clusterings, 1..256 pairs in random order and orientation, lists that nearly run empty):
    the kernels' arithmetic and the oracle agree on every pick, every fp32 score and the final running sums."""
    import itertools
    rng = np.random.RandomState(77)
    for _ in range(40):
        d, c, w = int(rng.randint(2, 24)), int(rng.randint(2, 40)), int(rng.randint(20, 500))
        picks = int(rng.randint(5, min(w, 120)))
        ids = rng.randint(0, c, size=(w, d))
        allp = list(itertools.combinations(range(d), 2))
        rng.shuffle(allp)
        pairs = [tuple(int(v) for v in (p if rng.rand() < 0.7 else p[::-1])) for p in allp[:int(rng.randint(1, min(len(allp), 256) + 1))]]
        want_pos, want_gain, want_sums = mo.greedy_mem_mi_pairs_c(ids, c, pairs, picks, return_sums=True)
        pos, gain, sums = _host_greedy(host_math, ids, c, pairs, picks)
        assert np.array_equal(pos, want_pos) and np.array_equal(gain, want_gain) and np.array_equal(sums, want_sums), (d, c, w)
