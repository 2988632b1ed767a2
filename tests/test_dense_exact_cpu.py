"""CPU checks of the bit-exact dense-MI arithmetic (csrc/mi_dense_exact_math.h, the header `acav_mi_dense_score_exact`
is built from), compiled for the host by tests/native/mi_dense_exact_host.cpp:

* its restatement of ATen's `.sum([2, 3])` order against torch itself on random rows (all cascade regimes);
* the dense MI of (table + one-hot), per (candidate, pair), against torch's evaluation of the reference expression
  (measures/mi.py:85-91) BIT FOR BIT, in the serial order and in the lane-by-lane decomposition the warp kernel uses;
* a free-running batch_mi selection driven by these scores reproduces the goldens written by the unmodified reference.
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
f32p, u32p, i64p = (ctypes.POINTER(t) for t in (ctypes.c_float, ctypes.c_uint32, ctypes.c_int64))


@pytest.fixture(scope="module")
def host(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    out = str(tmp_path_factory.mktemp("native") / "libdense_exact_host.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared",
                           os.path.join(HERE, "native", "mi_dense_exact_host.cpp"), "-o", out])
    lib = ctypes.CDLL(out)
    lib.host_dense_exact_row_sum.restype = ctypes.c_float
    lib.host_dense_exact_row_sum.argtypes = [f32p, ctypes.c_int64]
    lib.host_dense_exact_score.restype = None
    lib.host_dense_exact_score.argtypes = [u32p, u32p, u32p, u32p, ctypes.c_int32, ctypes.c_int32, i64p, ctypes.c_int64,
                                           f32p, f32p, ctypes.c_int32, f32p]
    return lib


def test_row_sum_order_matches_torch_sum_over_last_two_dims(host):
    rng = np.random.RandomState(1)
    for c in (1, 2, 3, 5, 6, 8, 11, 16, 20, 31, 32, 45, 64, 100, 128, 181, 256, 300, 512):
        x = (rng.standard_normal((3, 2, c, c)) * 10 ** rng.uniform(-2, 2, (3, 2, c, c))).astype(np.float32)
        want = torch.from_numpy(x).sum([2, 3]).numpy()
        for b in range(3):
            for p in range(2):
                row = np.ascontiguousarray(x[b, p]).reshape(-1)
                got = host.host_dense_exact_row_sum(row.ctypes.data_as(f32p), row.size)
                assert np.float32(got) == want[b, p], (c, b, p)


def _score(host, tables, cells, C, mode):
    from acav100m_b200.subset_selection.measures import tables as T
    N, a, b, n = tables
    P = N.shape[0]
    logs = T.log_table(int(n.max()) + 8).numpy()
    consts = T.dense_exact_constants(C)
    out = np.empty((cells.shape[0], P), dtype=np.float32)
    host.host_dense_exact_score(N.ctypes.data_as(u32p), a.ctypes.data_as(u32p), b.ctypes.data_as(u32p),
                                n.ctypes.data_as(u32p), P, C, cells.ctypes.data_as(i64p), cells.shape[0],
                                logs.ctypes.data_as(f32p), consts.ctypes.data_as(f32p), mode, out.ctypes.data_as(f32p))
    return out


@pytest.mark.parametrize("C,P,picked", [(6, 1, 0), (8, 3, 17), (20, 1, 200), (33, 2, 5), (64, 2, 3000), (256, 1, 40_000)])
def test_dense_mi_of_table_plus_one_hot_equals_torch_bits(host, C, P, picked):
    """Counts tables with `picked` samples (some cells / marginals still empty), 12 candidates: torch evaluates
    mi.py:90 on the [12, P, C, C] tensors exactly as the reference does; the header must return the same fp32 bits."""
    rng = np.random.RandomState(C * 7 + P)
    a_ids = torch.from_numpy(np.stack([rng.randint(0, C, size=picked + 12), rng.randint(0, C, size=picked + 12)], 1)
                             .repeat(1, 0)).long()
    assign = torch.cat([a_ids, a_ids.flip(1)], dim=1)[:, :max(2, P + 1)]           # enough columns for P pairs
    pairs = [(i, i + 1) for i in range(P)]
    Nt = torch.full((P, C, C), bo.EPS)
    cache = {"N": Nt, "a": Nt.sum(1), "b": Nt.sum(2)}
    cache["n"] = cache["a"].sum(-1)
    if picked:
        add = bo.sample_tables(assign, pairs, torch.arange(picked), C)
        for key in cache:
            cache[key] = cache[key] + add[key].sum(0)
    batch = torch.arange(picked, picked + 12)
    tabs = bo.sample_tables(assign, pairs, batch, C)
    want = bo.dense_mi({key: cache[key].unsqueeze(0) + tabs[key] for key in tabs}).numpy()        # [12, P]
    counts = np.zeros((P, C, C), dtype=np.uint32)
    for p, (u, v) in enumerate(pairs):
        np.add.at(counts[p], (assign[:picked, u].numpy(), assign[:picked, v].numpy()), 1)
    tables = (counts, counts.sum(1).astype(np.uint32), counts.sum(2).astype(np.uint32),
              np.full(P, picked, dtype=np.uint32))
    cells = np.ascontiguousarray(np.stack([assign[batch][:, [u, v]].numpy() for u, v in pairs], axis=1).astype(np.int64))
    for mode in (0, 1):
        got = _score(host, tables, cells, C, mode)
        assert np.array_equal(got, want), (mode, np.abs(got - want).max())


@pytest.mark.parametrize("name", sorted(gen_golden.BATCH_MI_CASES))
def test_free_running_batch_mi_with_header_scores_reproduces_reference_golden(host, golden_dir, name):
    """The reference's batch_mi loop (batch.py:93-165) with the scoring replaced by the header's evaluation, run freely
    from the golden's seed: S and GAIN must be the unmodified reference's.  (The GPU test does the same through
    acav_mi_dense_score_exact; this one needs no GPU.)"""
    from acav100m_b200.subset_selection.measures import tables as T
    g = dict(np.load(os.path.join(golden_dir, name + ".npz")))
    a = torch.from_numpy(g["assignments"].astype(np.int64))
    C, pairs = int(g["c"]), [tuple(p) for p in g["pairs"].tolist()]
    keep, V, subset = bool(g["keep_unselected"]), a.shape[0], int(g["subset"])
    P = len(pairs)
    counts = np.zeros((P, C, C), dtype=np.uint32)
    n_added = 0

    def add(ids):
        nonlocal n_added
        for p, (u, v) in enumerate(pairs):
            np.add.at(counts[p], (a[ids, u].numpy(), a[ids, v].numpy()), 1)
        n_added += len(ids)

    def score(batch):
        cells = np.ascontiguousarray(np.stack([a[batch][:, [u, v]].numpy() for u, v in pairs], axis=1).astype(np.int64))
        tabs = (counts, counts.sum(1).astype(np.uint32), counts.sum(2).astype(np.uint32), np.full(P, n_added, dtype=np.uint32))
        per_pair = torch.from_numpy(_score(host, tabs, cells, C, 1))
        return per_pair.mean(dim=-1)                               # torch's own mean over the pairs

    B = min(20, V - 1)
    k = bo.modify_k(4, B, subset, V, keep)
    add(torch.tensor([0]))
    cand = torch.arange(1, V)
    gen_golden.seed_all(int(g["seed"]))
    S, GAIN = [], []
    while len(S) < subset:
        cand = cand.index_select(0, torch.randperm(cand.shape[0]))
        batch = cand[:B]
        scores = score(batch)
        kk = k if scores.shape[0] >= B else int(np.floor(B / k * scores.shape[0]))
        top, ids = scores.topk(k=kk, dim=0)
        chosen = batch.index_select(0, ids)
        add(chosen)
        cand = cand[B:]
        if keep:
            u, c = torch.cat((batch, chosen)).unique(return_counts=True)
            cand = torch.cat((cand, u[c == 1]))
        S += chosen.tolist()
        GAIN += top.tolist()
    assert S[:subset] == g["S"].tolist()
    assert np.array_equal(np.array(GAIN, dtype=np.float32)[:len(g["GAIN"])], g["GAIN"].astype(np.float32)[:len(GAIN)])
