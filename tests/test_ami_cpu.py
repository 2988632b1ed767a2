"""CPU checks of the adjusted-MI measure `ami` (reference measures/mi.py:212-262) that need no GPU.

* the torch restatement (oracle/batch_mi_oracle.py) reproduces the goldens written by the unmodified reference bit for bit;
* the arithmetic header of the `ami` kernels (csrc/mi_ami_math.h), built for the host and walked through the kernels'
  data flow (row / column sums + four corrections per candidate), equals the dense fp64 evaluation of the reference's
  expressions to 1e-10, and the reference's own fp32 values to 1e-5 at the sizes it can run.
"""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest
import torch

from oracle import batch_mi_oracle as bo, gen_golden

HERE = os.path.dirname(os.path.abspath(__file__))
AMI = sorted(gen_golden.AMI_CASES)


def _golden(golden_dir, name):
    g = dict(np.load(os.path.join(golden_dir, name + ".npz")))
    order = g["candidate_order"].tolist()
    return g, g["assignments"].astype(np.int64), order, [tuple(p) for p in g["pairs"].tolist()]


@pytest.mark.parametrize("name", AMI)
def test_ami_restatement_reproduces_reference_bits(golden_dir, name):
    g, a, order, pairs = _golden(golden_dir, name)
    S, GAIN, _ = bo.greedy_dense_mi(a, int(g["c"]), pairs, order[1:], int(g["subset"]), [order[0]], measure="ami")
    assert S == g["S"].tolist()
    assert np.array_equal(np.array(GAIN), g["GAIN"])


@pytest.fixture(scope="module")
def host_ami(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    out = str(tmp_path_factory.mktemp("native") / "libmi_ami_math_host.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-o", out,
                           os.path.join(HERE, "native", "mi_ami_math_host.cpp")])
    lib = ctypes.CDLL(out)
    lib.host_ami_scores.restype = None
    lib.host_ami_scores.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32,
                                    ctypes.c_void_p]
    return lib


def _dense_truth(N, cells, average_method, dtype):
    """AMI of (table + candidate) for one pair, by the reference's dense expressions in `dtype`."""
    C = N.shape[0]
    tab = torch.full((1, C, C), bo.EPS, dtype=dtype) + torch.from_numpy(N.astype(np.float64)).to(dtype)[None]
    cache = {"N": tab, "a": tab.sum(dim=1), "b": tab.sum(dim=2)}
    cache["n"] = cache["a"].sum(dim=-1)
    a = torch.from_numpy(cells.astype(np.int64))
    tabs = bo.sample_tables(a, [(0, 1)], torch.arange(len(cells)), C)
    last = {k: cache[k].unsqueeze(0) + tabs[k].to(dtype) for k in tabs}
    return bo.dense_ami(last, average_method)[:, 0]


@pytest.mark.parametrize("C,n,method,seed", [(5, 40, "arithmetic", 1), (12, 400, "max", 2), (30, 3000, "min", 3),
                                             (7, 0, "arithmetic", 4), (64, 20000, "arithmetic", 5)])
def test_kernel_arithmetic_equals_dense_fp64_evaluation(host_ami, C, n, method, seed):
    rng = np.random.RandomState(seed)
    N = np.zeros((C, C), dtype=np.uint32)
    if n:
        picks = rng.randint(0, C, size=(n, 2))
        picks[:, 1] = np.where(rng.random_sample(n) < 0.6, (picks[:, 0] * 3 + 1) % C, picks[:, 1])
        np.add.at(N, (picks[:, 0], picks[:, 1]), 1)
    cells = np.ascontiguousarray(np.stack(np.meshgrid(np.arange(C), np.arange(C), indexing="ij"), -1).reshape(-1, 2),
                                 dtype=np.int32)
    out = np.zeros(len(cells), dtype=np.float64)
    host_ami.host_ami_scores(N.ctypes.data, C, cells.ctypes.data, len(cells), {"arithmetic": 0, "max": 1, "min": 2}[method],
                             out.ctypes.data)
    want = _dense_truth(N, cells, method, torch.float64).numpy()
    np.testing.assert_allclose(out, want, rtol=1e-9, atol=1e-11)
    if n <= 3000:                                                   # where fp32 lgamma differences still resolve
        want32 = _dense_truth(N, cells, method, torch.float32).numpy()
        np.testing.assert_allclose(out, want32, rtol=2e-5, atol=1e-6)


@pytest.mark.parametrize("name", AMI)
def test_kernel_arithmetic_follows_the_reference_run(host_ami, golden_dir, name):
    """The GPU test's replay (tests/test_zz_ami_gpu.py) with the host build standing in for the kernels: teacher-forced
    along the reference's own `ami` run, every candidate's score within 1e-5 of the reference's fp32 value and 1e-6 of
    the fp64 evaluation at every iteration."""
    g, a, order, pairs = _golden(golden_dir, name)
    C, subset, S_ref = int(g["c"]), int(g["subset"]), g["S"].tolist()
    _, _, ALL = bo.greedy_dense_mi(a, C, pairs, order[1:], subset, [order[0]], follow=S_ref[1:], measure="ami")
    _, _, ALL64 = bo.greedy_dense_mi(a, C, pairs, order[1:], subset, [order[0]], follow=S_ref[1:], measure="ami",
                                     dtype=torch.float64)
    tables = [np.zeros((C, C), dtype=np.uint32) for _ in pairs]
    for it, ((want, cand), (want64, _)) in enumerate(zip(ALL, ALL64)):
        per_pair = np.zeros((len(cand), len(pairs)), dtype=np.float32)
        for p, (c1, c2) in enumerate(pairs):
            cells = np.ascontiguousarray(a[cand.numpy()][:, [c1, c2]], dtype=np.int32)
            out = np.zeros(len(cells), dtype=np.float64)
            host_ami.host_ami_scores(tables[p].ctypes.data, C, cells.ctypes.data, len(cells), 0, out.ctypes.data)
            per_pair[:, p] = out.astype(np.float32)
        got = np.zeros(len(cand), dtype=np.float32)
        for p in range(len(pairs)):                                 # mi_dense_mean_kernel: pairs added in order
            got = (got + per_pair[:, p]).astype(np.float32)
        got = (got / np.float32(len(pairs))).astype(np.float32)
        np.testing.assert_allclose(got, want.numpy(), rtol=1e-5, atol=1e-6, err_msg="iteration %d" % it)
        np.testing.assert_allclose(got, want64.numpy(), rtol=1e-6, atol=1e-7, err_msg="iteration %d" % it)
        row = a[S_ref[1 + it]]
        for p, (c1, c2) in enumerate(pairs):
            tables[p][row[c1], row[c2]] += 1


@pytest.mark.parametrize("n_same", [1, 2, 37])
def test_singular_start_of_a_run_matches_the_reference(host_ami, n_same):
    """All picks so far in one cell: both entropies vanish, the denominator is a difference of 1e-14 terms, and the
    reference returns exactly 0 for a candidate joining that cell (its MI and EMI are bit-identical there), ~0 for one
    sharing its row or column, 1 for one on a fresh row and column."""
    C = 5
    N = np.zeros((C, C), dtype=np.uint32)
    N[2, 3] = n_same
    cells = np.ascontiguousarray(np.stack(np.meshgrid(np.arange(C), np.arange(C), indexing="ij"), -1).reshape(-1, 2),
                                 dtype=np.int32)
    out = np.zeros(len(cells), dtype=np.float64)
    host_ami.host_ami_scores(N.ctypes.data, C, cells.ctypes.data, len(cells), 0, out.ctypes.data)
    want32 = _dense_truth(N, cells, "arithmetic", torch.float32).numpy()
    np.testing.assert_allclose(out, want32, rtol=1e-5, atol=1e-6)
    assert out[2 * C + 3] == 0.0


def test_reference_fp32_noise_grows_with_the_table():
    """Why parity for `ami` is stated against the fp64 value: the reference's fp32 lgamma terms are ~n*log(n) each and
    their difference is O(1), so its relative error grows with n (1e-6 at n = 400, > 1e-4 by n = 10^5)."""
    rng = np.random.RandomState(8)
    errs = []
    for n in (400, 100_000):
        C = 8
        N = np.zeros((C, C), dtype=np.uint32)
        picks = rng.randint(0, C, size=(n, 2))
        np.add.at(N, (picks[:, 0], picks[:, 1]), 1)
        cells = np.array([[1, 2], [3, 3], [0, 7]], dtype=np.int32)
        d = _dense_truth(N, cells, "arithmetic", torch.float64).numpy()
        f = _dense_truth(N, cells, "arithmetic", torch.float32).numpy()
        errs.append(float(np.max(np.abs(f - d) / np.maximum(np.abs(d), 1e-12))))
    assert errs[0] < 1e-4 < errs[1], errs
