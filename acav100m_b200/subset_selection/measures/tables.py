"""Host-side constants that make the device arithmetic bit-identical to the reference's.

The reference evaluates ``x * x.log()`` with torch's CPU fp32 ``log`` on tensors that only ever hold
(a) exact integers >= 1 and (b) the "empty" values eps, C*eps, C*C*eps (``init_cache``,
measures/mi.py:32-39, :297-308).  CUDA's ``logf`` rounds differently from torch's CPU vector math
library in ~1e-5 of the arguments, enough to flip greedy picks, so the device never calls log():
it reads these tables, produced once on the host by the same torch operators.  This is setup, not
the hot loop: O(subset_size) work per run.
"""
import numpy as np
import torch

EPS = np.finfo('float64').eps       # measures/mi.py:25


def log_table(n):
    """fp32 log(k), k = 0..n-1, as torch's CPU kernel returns it (entry 0 is unused)."""
    t = torch.arange(0, max(int(n), 2), dtype=torch.float32).log()
    t[0] = 0.0
    return t


_DEVICE_TABLES = {}


def log_table_device(n, device):
    """`log_table(n)` resident on `device`, shared by every engine of the process: the table is a constant of the
    torch build (not of the data), so the longest one made so far is kept per device and reused (the chunked runner
    and multi-layer jobs create many engines; 2^24 entries take ~30 ms to produce and upload)."""
    key = str(device)
    have = _DEVICE_TABLES.get(key)
    if have is None or have.numel() < n:
        have = log_table(n).to(device)
        _DEVICE_TABLES[key] = have
    return have


def empty_table_constants(C):
    """{fN0, fa0, n0, NlogN0, aloga0, blogb0} of an empty C x C table (one clustering pair)."""
    N = torch.full((1, C, C), EPS)
    a = N.sum(dim=1)
    b = N.sum(dim=2)
    n = a.sum(dim=-1)
    xlogx = lambda v: v * v.log()
    consts = [xlogx(N[0, 0, 0]), xlogx(a[0, 0]), n[0], xlogx(N).sum([-1, -2])[0], xlogx(a).sum(-1)[0],
              xlogx(b).sum(-1)[0]]
    assert float(N[0, 0, 0] + 1) == 1.0 and float(a[0, 0] + 1) == 1.0 and float(n[0] + 1) == 1.0
    return np.array([float(c) for c in consts], dtype=np.float32)


def pair_table_constants(P, C):
    """fp32 [P, 6]: per clustering pair {fN0, fa0, n0, NlogN0, aloga0, blogb0} of the empty tables, as the reference's
    ``init_cache`` (measures/mi.py:32-39, :297-308) produces them on its [P, C, C] tensors.  With two or more pairs every
    output element of those reductions is summed serially by one thread (torch parallelises over outputs, not inside a
    row), so the values do not depend on P; they are computed on a [2, C, C] tensor and repeated (checked against the
    full shape in tests/test_mi_pairs_math_cpu.py).  One pair goes through `empty_table_constants`."""
    if P == 1:
        return empty_table_constants(C).reshape(1, 6)
    N = torch.full((2, C, C), EPS)
    a = N.sum(dim=1)
    b = N.sum(dim=2)
    n = a.sum(dim=-1)
    xlogx = lambda v: v * v.log()
    row = [xlogx(N[0, 0, 0]), xlogx(a[0, 0]), n[0], xlogx(N).sum([-1, -2])[0], xlogx(a).sum(-1)[0], xlogx(b).sum(-1)[0]]
    assert float(N[0, 0, 0] + 1) == 1.0 and float(a[0, 0] + 1) == 1.0 and float(n[0] + 1) == 1.0
    return np.tile(np.array([float(c) for c in row], dtype=np.float32), (P, 1))


def dense_exact_constants(C):
    """fp32 {eps, a0, b0, log eps, log a0, log b0}: an empty cell, an empty column marginal (``N.sum(dim=1)``) and an empty
    row marginal (``N.sum(dim=2)``) of ``init_cache`` (measures/mi.py:32-39) and their torch-CPU logs, for the bit-exact
    dense scorer (acav_mi_dense_score_exact).  Computed on a [2, C, C] tensor: with two or more outputs torch sums every
    output serially, so the values do not depend on the number of pairs."""
    N = torch.full((2, C, C), EPS)
    a = N.sum(dim=1)
    b = N.sum(dim=2)
    vals = torch.stack([N[0, 0, 0], a[0, 0], b[0, 0]])
    assert bool((a == a[0, 0]).all()) and bool((b == b[0, 0]).all())
    assert float(vals[0] + 1) == 1.0 and float(vals[1] + 1) == 1.0 and float(vals[2] + 1) == 1.0
    assert float(a.sum(dim=-1)[0] + 1) == 1.0                      # n0 = C*C*eps is absorbed by the first sample
    return np.concatenate([vals.numpy(), vals.log().numpy()]).astype(np.float32)
