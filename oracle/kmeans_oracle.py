"""CPU restatement of the reference's mini-batch SGD k-means operator.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) -- the product never imports this file.

Follows ``clustering/code/sgd_clustering.py`` of the reference (file:line cited per function).  The
arithmetic of that file lives in un-vendored third-party code -- ``torch`` (matmul / norm / min /
rand, reference pins 1.6.0) and ``torch-scatter==2.0.5`` (``scatter_add``) -- so the restatement
calls the same torch CPU operators in the same order: on one machine it is bit-identical to the
reference run through ``oracle/ref_shims.py``, which is how it is pinned
(``tests/golden/kmeans_*.npz`` written by ``oracle/gen_golden.py``,
checked in ``tests/test_oracle_golden.py``).

``assign_truth_f64`` is NOT a restatement: it is the fp64 evaluation of the same distance formula,
used by the parity tests to decide whether a row whose fp32 argmin differs between two fp32
implementations (MKL vs tensor cores) is a genuine near-tie.
"""
import dataclasses
import math

import numpy as np
import torch


@dataclasses.dataclass
class SgdKMeansState:
    """Operator state, reference ``KMeans.__init__`` sgd_clustering.py:24-32."""
    centers: torch.Tensor            # [k, d] fp32
    counts: torch.Tensor             # [k] fp32 (integers stored as floats, :26)
    count: int = 0                   # samples seen, all ranks (:27, :128)
    lr: float = 1e-2
    initial_rounds: int = 10
    reinit: tuple = (.7, 5.0)
    fallback: int = 0                # number of lr fallbacks taken (:119)
    sequential: bool = False         # slow sequential update branch (:103-109), off by default (:32)

    def clone(self):
        return dataclasses.replace(self, centers=self.centers.clone(), counts=self.counts.clone())


def new_state(d, k, lr=1e-2, initial_rounds=10, reinit=(.7, 5.0)):
    """sgd_clustering.py:24-31 -- centers drawn from torch's global CPU generator."""
    return SgdKMeansState(centers=torch.rand(k, d) * 1e-5, counts=torch.zeros(k), count=0,
                          lr=lr, initial_rounds=initial_rounds, reinit=tuple(reinit))


def in_warmup(state):
    """sgd_clustering.py:67 -- random assignment until initial_rounds*k samples were seen."""
    return state.count < state.initial_rounds * state.centers.shape[0]


def underused_mask(state):
    """sgd_clustering.py:76-77 -- centroids whose distances get divided by r."""
    k = state.centers.shape[0]
    p, _ = state.reinit
    return state.counts < (state.count / k) ** p


def distances_fp32(state, batch):
    """sgd_clustering.py:72-77 -- [k, b] fp32 squared distances incl. the re-init scaling."""
    dist = -2 * torch.matmul(state.centers, batch.T)
    dist += (torch.norm(batch, dim=1) ** 2)[None, :]
    dist += (torch.norm(state.centers, dim=1) ** 2)[:, None]
    _, r = state.reinit
    dist[underused_mask(state), :] /= r
    return dist


def assign(state, batch, warmup_noise=None):
    """``KMeans.calc_best`` sgd_clustering.py:63-79 -> (best int64[b], mean min distance float).

    `warmup_noise` ([k, b] fp32) replaces the ``torch.rand(k, b)`` draw of :68 when given, so that a
    caller can feed the same noise to two implementations; when None it is drawn from the global
    CPU generator exactly like the reference.
    """
    k = state.centers.shape[0]
    b = len(batch)
    if in_warmup(state):
        dist = torch.rand(k, b) if warmup_noise is None else warmup_noise
    else:
        dist = distances_fp32(state, batch)
    mind, best = dist.min(axis=0)
    return best, mind.mean().item()


def effective_lr(lr, max_count):
    """sgd_clustering.py:116-119 -- python-float (double) arithmetic on the host."""
    if max_count * lr >= 1.0:
        return 0.5 / max_count, True
    return lr, False


def sgd_step_sequential(state, batch, warmup_noise=None, best=None):
    """``KMeans.add`` with ``sequential=True`` (sgd_clustering.py:103-109): rows are applied one at a time, in batch
    order, each to its centroid: ``c *= 1 - lr; c += lr * x; counts += 1``.  No lr fallback, no decay by the histogram."""
    lr = state.lr(state.count) if callable(state.lr) else state.lr
    if best is None:
        best, mean_dist = assign(state, batch, warmup_noise)
    else:
        mean_dist = float("nan")
    for i, j in enumerate(best):
        state.centers[j] *= 1 - lr                                                   # :107
        state.centers[j] += lr * batch[i]                                            # :108
        state.counts[j] += 1                                                         # :109
    state.count += len(batch)                                                        # :128
    return best, mean_dist


def sgd_step(state, batch, warmup_noise=None, best=None):
    """``KMeans.add`` sgd_clustering.py:94-129, single process (fast parallel update branch; the sequential branch
    when ``state.sequential``).

    Mutates `state`; returns (best, mean distance).  `best` may be forced (used to test the update
    in isolation).
    """
    if state.sequential:
        return sgd_step_sequential(state, batch, warmup_noise, best)
    k, d = state.centers.shape
    lr = state.lr(state.count) if callable(state.lr) else state.lr
    if best is None:
        best, mean_dist = assign(state, batch, warmup_noise)
    else:
        mean_dist = float("nan")
    ones = torch.ones(len(batch), dtype=torch.float)
    counts = torch.zeros(k, dtype=torch.float).scatter_add_(0, best, ones)          # :113
    lr, fell_back = effective_lr(lr, counts.max().item())                            # :116-119
    state.fallback += int(fell_back)
    state.counts += counts                                                           # :120
    state.centers *= (1. - counts * lr)[:, None]                                     # :121
    deltas = torch.zeros_like(state.centers)                                         # :122
    deltas.scatter_add_(0, best[:, None].expand(-1, d), batch * lr)                  # :123
    state.centers = state.centers + deltas                                           # :127
    state.count += len(batch)                                                        # :128
    return best, mean_dist


def sgd_step_world(state, rank_batches, warmup_noises=None):
    """``KMeans.add`` with is_distributed (sgd_clustering.py:96-97,114-115,125-126): every rank
    assigns its own slice against the replicated centers, histograms and deltas are summed over
    ranks (rank order = NCCL-order stand-in), `count` advances by the GLOBAL batch.
    Returns list of per-rank (best, mean distance)."""
    k, d = state.centers.shape
    lr = state.lr(state.count) if callable(state.lr) else state.lr
    outs = []
    for r, xb in enumerate(rank_batches):
        noise = None if warmup_noises is None else warmup_noises[r]
        outs.append(assign(state, xb, noise))
    counts = torch.zeros(k, dtype=torch.float)
    for (best, _), xb in zip(outs, rank_batches):
        counts += torch.zeros(k, dtype=torch.float).scatter_add_(0, best, torch.ones(len(xb)))
    lr, fell_back = effective_lr(lr, counts.max().item())
    state.fallback += int(fell_back)
    state.counts += counts
    state.centers *= (1. - counts * lr)[:, None]
    deltas = torch.zeros_like(state.centers)
    for (best, _), xb in zip(outs, rank_batches):
        local = torch.zeros_like(state.centers)
        local.scatter_add_(0, best[:, None].expand(-1, d), xb * lr)
        deltas += local
    state.centers = state.centers + deltas
    state.count += sum(len(xb) for xb in rank_batches)
    return outs


def epoch_lr(epoch):
    """run_clustering.py:168 -- lr schedule set by the driver at the start of every epoch."""
    return 0.1 ** (2 + epoch // 5)


def effective_epochs(epochs, num_gpus):
    """run_clustering.py:146."""
    return math.ceil(epochs / num_gpus)


def train(state, batches_per_epoch, epochs, pre_epochs=0):
    """Host loop of ``train_clusters`` run_clustering.py:164-176 for one clustering."""
    dists = []
    for epoch in range(pre_epochs, epochs + pre_epochs):
        state.lr = epoch_lr(epoch)
        for xb in batches_per_epoch(epoch):
            dists.append(sgd_step(state, xb)[1])
    return dists


# ----------------------------------------------------------------------------------------------
# fp64 truth (attribution of fp32 near-ties; not a restatement of reference code)
# ----------------------------------------------------------------------------------------------

def assign_truth_f64(centers, batch, underused=None, r=5.0):
    """fp64 evaluation of the distance formula of sgd_clustering.py:72-77.

    Returns (best int64[b], d1 float64[b], d2 float64[b]) -- smallest and second-smallest distance
    per row, first index on exact ties.
    """
    c = np.asarray(centers, dtype=np.float64)
    x = np.asarray(batch, dtype=np.float64)
    dist = -2.0 * (c @ x.T)
    dist += (x * x).sum(1)[None, :]
    dist += (c * c).sum(1)[:, None]
    if underused is not None:
        dist[np.asarray(underused, dtype=bool), :] /= r
    best = dist.argmin(axis=0)
    if dist.shape[0] > 1:
        part = np.partition(dist, 1, axis=0)
        d1, d2 = part[0], part[1]
    else:
        d1 = dist[0]
        d2 = np.full_like(d1, np.inf)
    return best.astype(np.int64), d1, d2


def fp32_ambiguity_band(centers, batch, ulps=64.0):
    """Per-row width below which two fp32 evaluations of the formula may legitimately disagree.

    The three fp32 roundings of (-2*dot + |x|^2) + |c|^2 and the sgemm accumulation error are all
    bounded by a small multiple of eps32 * (|x|^2 + |c|^2 + 2|x||c|) = eps32 * (|x| + |c|max)^2.
    `ulps` is that multiple (generous: d-term accumulation grows like sqrt(d) in practice).
    """
    c = np.asarray(centers, dtype=np.float64)
    x = np.asarray(batch, dtype=np.float64)
    xn = np.sqrt((x * x).sum(1))
    cn = np.sqrt((c * c).sum(1)).max()
    return ulps * np.finfo(np.float32).eps * (xn + cn) ** 2


def sequential_scatter_sum(values, index, k):
    """Plain row-order loop -- the definition torch-scatter 2.0.5's CPU kernel implements
    (``for i: out[index[i]] += src[i]``); used to check that torch's ``scatter_add_`` and the CUDA
    segmented sum both reproduce strict row order."""
    out = np.zeros((k,) + values.shape[1:], dtype=values.dtype)
    for i, j in enumerate(np.asarray(index)):
        out[j] += values[i]
    return out
