"""CPU restatement of the reference's exact greedy mutual-information subset selection (``mem_mi``).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) -- the product never imports this file.

Follows ``subset_selection/code/measures/mi.py`` (``EfficientMI`` :14-192, ``EfficientMemMI``
:284-412) and ``run_greedy.py:9-54`` / ``pairing.py`` of the reference; file:line cited per function.
All table arithmetic is fp32 with torch CPU operators (``log`` comes from torch's CPU vector math
library, which is not bit-identical to numpy's or CUDA's), so the restatement calls the same
operators and is pinned bit-for-bit against the reference itself: ``tests/golden/mi_*.npz`` written
by ``oracle/gen_golden.py``, checked in ``tests/test_oracle_golden.py``.

Two forms are provided:

* ``greedy_mem_mi``      -- torch restatement, any number of clustering pairs P, O(W*P) per iteration;
* ``greedy_mem_mi_c``    -- the plain-C restatement in ``oracle/mi_oracle.c`` (P = 1), either as the
                           literal per-candidate scan or bucketed by contingency-table cell
                           (same picks, O(C*C) per iteration) so that full-size configs finish on CPU;
* ``greedy_mem_mi_pairs_c`` -- plain C, any P: literal scan with the mean over pairs added in the order of torch's
                           CPU reduction kernel (restated in mi_oracle.c, pinned against torch and the goldens).
"""
import ctypes
import itertools
import os
import subprocess
from collections import defaultdict

import numpy as np
import torch

EPS = np.finfo('float64').eps          # mi.py:25 (becomes 2**-52 in fp32 tables, mi.py:35)
_HERE = os.path.dirname(os.path.abspath(__file__))


# ----------------------------------------------------------------------------------------------
# pairing.py
# ----------------------------------------------------------------------------------------------

def cluster_pairing(keys, kind):
    """pairing.py:5-41 -- which clustering columns form contingency tables."""
    kind = kind.lower()
    if kind == 'combination':                      # :16-20
        return list(itertools.combinations(range(len(keys)), 2))
    groups = defaultdict(list)
    if kind == 'bipartite':                        # :23-30, grouped by key[0]
        for i, key in enumerate(keys):
            groups[key[0]].append(i)
        return list(itertools.product(*groups.values()))
    if kind == 'diagonal':                         # :33-41, grouped by key[1]
        for i, key in enumerate(keys):
            groups[key[1]].append(i)
        return list(groups.values())
    raise AssertionError(f"invalid cluster pairing type: {kind}")


# ----------------------------------------------------------------------------------------------
# table state
# ----------------------------------------------------------------------------------------------

def init_table(P, C):
    """``EfficientMI.init_cache`` mi.py:32-39 + ``EfficientMemMI.init_cache`` :297-308."""
    N = torch.full((P, C, C), EPS)                 # fp32 (torch.full with a python float)
    a = N.sum(dim=1)                               # [P, C] indexed by c2 (column marginal)
    b = N.sum(dim=2)                               # [P, C] indexed by c1 (row marginal)
    n = a.sum(dim=-1)                              # [P]
    return {
        'N': N, 'a': a, 'b': b, 'n': n,
        'NlogN': (N * N.log()).sum([-1, -2]),
        'aloga': (a * a.log()).sum(-1),
        'blogb': (b * b.log()).sum(-1),
    }


def candidate_cells(assignments, pairs, candidates):
    """``EfficientMemMI.calc_N`` mi.py:285-291 -> int64 [W, P, 2] of (c1, c2) per pair."""
    rows = torch.from_numpy(np.asarray(assignments)).to(torch.long)
    rows = rows.index_select(0, torch.as_tensor(list(candidates), dtype=torch.long))
    pair_ids = torch.as_tensor(list(pairs), dtype=torch.long)       # [P, 2]
    return rows[:, pair_ids]                                        # [W, P, 2]


def _xlogx(v):
    return v * v.log()                             # mi.py:335-337


def candidate_scores(tab, cells):
    """``get_last`` mi.py:322-333 + ``calc_MI`` :368-381 for every remaining candidate.

    Returns (scores [W, P], NlogN' [W, P], aloga' [W, P], blogb' [W, P])."""
    P = tab['N'].shape[0]
    p = torch.arange(P)[None, :]
    c1, c2 = cells[:, :, 0], cells[:, :, 1]
    x = tab['N'][p, c1, c2]                        # get_last_N :342-348
    y = tab['a'][p, c2]                            # get_last_ab(dim=1) :326
    z = tab['b'][p, c1]                            # get_last_ab(dim=0) :327
    NlogN = tab['NlogN'][None] - _xlogx(x) + _xlogx(x + 1)          # update_nlogn :339-340
    aloga = tab['aloga'][None] - _xlogx(y) + _xlogx(y + 1)
    blogb = tab['blogb'][None] - _xlogx(z) + _xlogx(z + 1)
    n = (tab['n'] + 1)[None]                       # :332
    scores = ((NlogN / n) + (-aloga / n) + (-blogb / n)) + n.log()  # :375-380
    return scores, NlogN, aloga, blogb


def apply_pick(tab, cell, NlogN, aloga, blogb):
    """``update_cache`` mi.py:383-389 + ``update_mats`` :401-406 for the winning candidate."""
    tab['NlogN'], tab['aloga'], tab['blogb'] = NlogN.clone(), aloga.clone(), blogb.clone()
    for p in range(cell.shape[0]):
        c1, c2 = int(cell[p, 0]), int(cell[p, 1])
        tab['N'][p, c1, c2] += 1
        tab['a'][p, c2] += 1
        tab['b'][p, c1] += 1
    tab['n'] += 1


def greedy_mem_mi(assignments, ncentroids, pairs, candidates, subset_size, start_indices):
    """``EfficientMI.run_greedy`` mi.py:150-192 with the ``mem_mi`` measure.

    Returns (S, GAIN): S begins with `start_indices` (which are NOT added to the table) and gains
    ``subset_size - 1 - len(start_indices)`` picks; ties go to the earliest remaining candidate
    (``max(dim=0)`` mi.py:79)."""
    tab = init_table(len(pairs), ncentroids)
    ids = torch.as_tensor(list(candidates), dtype=torch.long)
    cells = candidate_cells(assignments, pairs, candidates)
    S, GAIN = list(start_indices), []
    for _ in range(len(start_indices), subset_size - 1):
        scores, NlogN, aloga, blogb = candidate_scores(tab, cells)
        score, idx = scores.mean(dim=-1).max(dim=0)                 # calc_score :76-80
        idx = idx.item()
        S.append(ids[idx].item())
        GAIN.append(score.item())
        apply_pick(tab, cells[idx], NlogN[idx], aloga[idx], blogb[idx])
        keep = torch.ones(len(ids), dtype=torch.bool)               # _remove_idx :124-125
        keep[idx] = False
        ids, cells = ids[keep], cells[keep]
    return S, GAIN


def run_greedy_driver(assignments, subset_size=None, subset_ratio=0.2, pairing='combination',
                      clustering_types=None, shuffle_candidates=False, rng=None):
    """``_run_greedy`` run_greedy.py:9-54 for measure_name='mem_mi' -> (S, GAIN)."""
    assignments = np.asarray(assignments)
    ncentroids = int(assignments.max()) + 1                          # :20
    V = assignments.shape[0]
    if subset_size is None:
        subset_size = round(subset_ratio * V)                        # :22-23
    if clustering_types is None:
        clustering_types = [('m%d' % i, 'layer') for i in range(assignments.shape[1])]
    pairs = cluster_pairing(clustering_types, pairing)
    candidates = list(range(V))
    if shuffle_candidates:
        (rng or __import__('random')).shuffle(candidates)            # :37-40
    start, candidates = [candidates[0]], candidates[1:]              # :43-44
    return greedy_mem_mi(assignments, ncentroids, pairs, candidates, subset_size, start)


# ----------------------------------------------------------------------------------------------
# plain-C restatement (P = 1)
# ----------------------------------------------------------------------------------------------

def log_table(n_max):
    """fp32 ``log(k)`` for k = 0..n_max as torch's CPU kernel returns it (entry 0 unused).

    The reference evaluates ``x.log()`` on fp32 tensors holding exact integers (counts >= 1), so a
    table of torch's own results reproduces its bits; verified position-independent in
    tests/test_oracle_golden.py."""
    t = torch.arange(0, n_max + 1, dtype=torch.float32)
    out = t.log()
    out[0] = 0.0
    return out.numpy()


def zero_count_constants(C):
    """Values the fp32 tables hold for a count of zero, and their x*log(x) (mi.py:32-39,297-308)."""
    tab = init_table(1, C)
    n0 = tab['N'][0, 0, 0]
    a0 = tab['a'][0, 0]
    return {
        'fN0': float(_xlogx(n0)), 'fa0': float(_xlogx(a0)),
        'n0': float(tab['n'][0]), 'NlogN0': float(tab['NlogN'][0]),
        'aloga0': float(tab['aloga'][0]), 'blogb0': float(tab['blogb'][0]),
        'N0_plus_1': float(n0 + 1), 'a0_plus_1': float(a0 + 1),
    }


_LIB = None


def c_library_path():
    return os.path.join(_HERE, 'libmi_oracle.so')


def build_c_oracle(force=False):
    """gcc -O2 -ffp-contract=off: no FMA contraction, IEEE fp32 throughout."""
    src, out = os.path.join(_HERE, 'mi_oracle.c'), c_library_path()
    if force or not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
        subprocess.check_call(['gcc', '-O2', '-ffp-contract=off', '-fno-fast-math', '-fPIC', '-shared',
                               '-fopenmp', '-o', out, src, '-lm'])
    return out


def _lib():
    global _LIB
    if _LIB is None:
        lib = ctypes.CDLL(build_c_oracle())
        f32p = ctypes.POINTER(ctypes.c_float)
        i32p = ctypes.POINTER(ctypes.c_int32)
        i64p = ctypes.POINTER(ctypes.c_int64)
        for name in ('mi_oracle_greedy_scan', 'mi_oracle_greedy_bucketed'):
            fn = getattr(lib, name)
            fn.restype = ctypes.c_int64
            fn.argtypes = [i32p, i32p, ctypes.c_int64, ctypes.c_int32, f32p, ctypes.c_int64,
                           f32p, ctypes.c_int64, i64p, f32p]
        lib.mi_oracle_aten_row_mean.restype = ctypes.c_float
        lib.mi_oracle_aten_row_mean.argtypes = [f32p, ctypes.c_int64]
        lib.mi_oracle_greedy_pairs.restype = ctypes.c_int64
        lib.mi_oracle_greedy_pairs.argtypes = [i32p, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, i32p,
                                               ctypes.c_int32, f32p, ctypes.c_int64, f32p, ctypes.c_int64, i64p,
                                               f32p, f32p]
        lib.mi_oracle_scan_once.restype = ctypes.c_double
        lib.mi_oracle_scan_once.argtypes = [i32p, i32p, ctypes.c_int64, ctypes.c_int32, f32p,
                                            ctypes.c_int64, f32p, ctypes.c_int32, ctypes.c_int32]
        _LIB = lib
    return _LIB


def _consts_array(C):
    z = zero_count_constants(C)
    return np.array([z['fN0'], z['fa0'], z['n0'], z['NlogN0'], z['aloga0'], z['blogb0']],
                    dtype=np.float32)


def greedy_mem_mi_c(c1, c2, C, n_picks, bucketed=True):
    """P = 1 exact greedy over candidates 0..W-1 given as cell coordinates (c1[w], c2[w]).

    Returns (positions int64[n_picks] into the candidate list, gains fp32[n_picks])."""
    c1 = np.ascontiguousarray(c1, dtype=np.int32)
    c2 = np.ascontiguousarray(c2, dtype=np.int32)
    W = len(c1)
    n_picks = int(min(n_picks, W))
    logs = log_table(n_picks + 2)
    consts = _consts_array(C)
    pos = np.zeros(n_picks, dtype=np.int64)
    gain = np.zeros(n_picks, dtype=np.float32)
    fn = _lib().mi_oracle_greedy_bucketed if bucketed else _lib().mi_oracle_greedy_scan
    as_p = lambda a, t: a.ctypes.data_as(ctypes.POINTER(t))
    done = fn(as_p(c1, ctypes.c_int32), as_p(c2, ctypes.c_int32), W, C,
              as_p(logs, ctypes.c_float), len(logs), as_p(consts, ctypes.c_float), n_picks,
              as_p(pos, ctypes.c_int64), as_p(gain, ctypes.c_float))
    assert done == n_picks, (done, n_picks)
    return pos, gain


def pair_constants(P, C):
    """Per pair {fN0, fa0, n0, NlogN0, aloga0, blogb0} of the empty tables, computed on the [P, C, C] tensors the
    reference builds (mi.py:32-39, :297-308): the fp32 sums over the empty tables depend on the tensor shape."""
    tab = init_table(P, C)
    out = np.zeros((P, 6), dtype=np.float32)
    for p in range(P):
        out[p] = [float(_xlogx(tab['N'][p, 0, 0])), float(_xlogx(tab['a'][p, 0])), float(tab['n'][p]),
                  float(tab['NlogN'][p]), float(tab['aloga'][p]), float(tab['blogb'][p])]
        assert float(tab['N'][p, 0, 0] + 1) == 1.0 and float(tab['a'][p, 0] + 1) == 1.0 and float(tab['n'][p] + 1) == 1.0
    return out


def aten_row_mean(row):
    """`row.mean()` of a contiguous fp32 row as torch's CPU inner-dimension reduction adds it (mi_oracle.c)."""
    row = np.ascontiguousarray(row, dtype=np.float32)
    return float(_lib().mi_oracle_aten_row_mean(row.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), len(row)))


def greedy_mem_mi_pairs_c(ids, C, pairs, n_picks, return_sums=False):
    """Exact greedy over P clustering pairs (literal scan, plain C).  ids: int [W, D] cluster ids of the candidates in
    list order; pairs: [(col1, col2)] * P.  Returns (positions int64[n], gains fp32[n]) (+ final sums [P, 4])."""
    ids = np.ascontiguousarray(ids, dtype=np.int32)
    W, D = ids.shape
    pr = np.ascontiguousarray(np.asarray(pairs, dtype=np.int32).reshape(-1, 2))
    P = len(pr)
    n_picks = int(min(n_picks, W))
    logs = log_table(n_picks + 2)
    consts = np.ascontiguousarray(pair_constants(P, C))
    pos = np.zeros(n_picks, dtype=np.int64)
    gain = np.zeros(n_picks, dtype=np.float32)
    sums = np.zeros((P, 4), dtype=np.float32)
    as_p = lambda a, t: a.ctypes.data_as(ctypes.POINTER(t))
    done = _lib().mi_oracle_greedy_pairs(as_p(ids, ctypes.c_int32), W, D, C, as_p(pr, ctypes.c_int32), P,
                                         as_p(logs, ctypes.c_float), len(logs), as_p(consts, ctypes.c_float), n_picks,
                                         as_p(pos, ctypes.c_int64), as_p(gain, ctypes.c_float),
                                         as_p(sums, ctypes.c_float))
    assert done == n_picks, (done, n_picks)
    return (pos, gain, sums) if return_sums else (pos, gain)


def scan_once_seconds(c1, c2, C, repeats, threads):
    """Time `repeats` full candidate scans of the literal C restatement (cpu_baseline leg)."""
    c1 = np.ascontiguousarray(c1, dtype=np.int32)
    c2 = np.ascontiguousarray(c2, dtype=np.int32)
    logs = log_table(repeats + 4)
    consts = _consts_array(C)
    as_p = lambda a, t: a.ctypes.data_as(ctypes.POINTER(t))
    return _lib().mi_oracle_scan_once(as_p(c1, ctypes.c_int32), as_p(c2, ctypes.c_int32), len(c1), C,
                                      as_p(logs, ctypes.c_float), len(logs),
                                      as_p(consts, ctypes.c_float), repeats, threads)


def greedy_mem_mi_via_c(assignments, subset_size, pair=(0, 1), bucketed=True):
    """Driver-level convenience (run_greedy.py:9-54, shuffle off, P = 1) on top of the C oracle."""
    assignments = np.asarray(assignments)
    C = int(assignments.max()) + 1
    cand = assignments[1:]
    pos, gain = greedy_mem_mi_c(cand[:, pair[0]], cand[:, pair[1]], C, max(subset_size - 2, 0),
                                bucketed=bucketed)
    return [0] + [int(p) + 1 for p in pos], gain.astype(np.float64).tolist()
