"""``EfficientMI`` -- the reference's dense greedy measure ``mi`` on a B200 (toy-size cross-check measure).

Mirror of ``subset_selection/code/measures/mi.py:14-209``: same constructor keywords, ``init(pairs,
candidates)``, ``add_samples``, ``run_greedy -> (S, GAIN, timelapse, LOOKUPS)``.  Per iteration the reference
materialises the dense table of every remaining candidate (``W x P x C x C`` floats, :47-59), evaluates
``sum_ij N/n (log N + log n - log a - log b)`` on each (:85-91), averages over the clustering pairs and takes
the first maximum (:76-80).  Here the candidates' cells stay on the device and every iteration is one
``acav_mi_dense_score`` call over all remaining candidates -- O(W*P) from the running sums of the table
instead of O(W*P*C*C) logs -- plus one ``acav_mi_dense_add`` for the winner.

Parity: with ``exact=True`` (default) every candidate is scored by ``acav_mi_dense_score_exact`` -- the reference's
own O(W*P*C*C) dense evaluation with its five fp32 roundings per cell, summed in the order of torch's CPU reduction
kernel -- so the scores carry the reference's bits and a free run selects the reference's indices
(``tests/test_dense_mi_gpu.py``).  ``exact=False`` scores in O(W*P) from fp64 running sums (``acav_mi_dense_score``):
~1e-6 relative agreement, picks identical except where two different cells are closer than the noise of the
reference's dense sum.
"""
import time

import numpy as np
import torch

from ... import _lib
from . import tables


def score_exact(measure, cells, w, scores):
    """``acav_mi_dense_score_exact`` for a measure object holding ``_engine``, ``_n_added``, ``ncentroids``, ``device``:
    the log table must cover every count the tables can hold plus the candidate's sample."""
    logs = tables.log_table_device(measure._n_added + 4, measure.device)
    consts = tables.dense_exact_constants(measure.ncentroids)
    _lib.call("acav_mi_dense_score_exact", measure._engine, _lib.ptr(cells), w, _lib.ptr(logs), logs.numel(),
              consts.ctypes.data_as(_lib.c_vp), _lib.ptr(scores), None, _lib.stream_ptr(measure.device))


class EfficientMI:
    def __init__(self, assignments, measure_type='mutual_info', average_method='arithmetic',
                 ncentroids=20, device=None, exact=True, **kwargs):
        self.average_method = average_method.lower()
        self.ncentroids = int(ncentroids)
        self.assignments = torch.from_numpy(np.asarray(assignments)).to(torch.long)      # V x D (mi.py:24)
        self.eps = tables.EPS
        self.device = _lib.require_cuda(device if device not in (None, 'cpu', 'cuda') else None)
        self.exact = exact
        self._engine = None
        self._n_added = 0

    def init(self, clustering_combinations, candidates):
        """mi.py:27-30."""
        self._n_added = 0
        self.combinations = [tuple(p) for p in clustering_combinations]
        self._pair_ids = torch.as_tensor(self.combinations, dtype=torch.long)            # P x 2
        self._release()
        handle = _lib.c_vp()
        with torch.cuda.device(self.device):
            _lib.call("acav_mi_dense_create", _lib.ctypes.byref(handle), len(self.combinations), self.ncentroids,
                      _lib.stream_ptr(self.device))
        self._engine = handle
        self.init_candidates(candidates)

    def init_candidates(self, candidates):
        """mi.py:47-59 -- the candidates' cells per pair (instead of their dense one-hot tables)."""
        self.candidate_ids = torch.as_tensor(np.asarray(candidates, dtype=np.int64))
        self._cand_cells = self._cells(self.candidate_ids)                               # [W, P, 2] on the device

    def _release(self):
        if self._engine is not None:
            _lib.load().acav_mi_dense_destroy(self._engine)
            self._engine = None

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    def _cells(self, ids):
        rows = self.assignments.index_select(0, ids)                                     # get_assignments mi.py:41-45
        cells = rows[:, self._pair_ids]                                                  # [m, P, 2]
        if cells.numel() and (int(cells.min()) < 0 or int(cells.max()) >= self.ncentroids):
            raise ValueError("cluster ids must lie in [0, ncentroids)")
        return cells.contiguous().to(self.device)

    def add_samples(self, ids):
        """mi.py:145-148."""
        cells = self._cells(torch.as_tensor(ids, dtype=torch.long))
        self._add_cells(cells)

    def _add_cells(self, cells):
        self._n_added += int(cells.shape[0])
        with torch.cuda.device(self.device):
            _lib.call("acav_mi_dense_add", self._engine, _lib.ptr(cells), cells.shape[0], _lib.stream_ptr(self.device))

    def score_candidates(self):
        """``get_last`` + ``calc_MI`` + ``mean(dim=-1)`` (mi.py:76-98) over all remaining candidates -> fp32 [W]."""
        w = self._cand_cells.shape[0]
        scores = torch.empty(w, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            self._score(self._cand_cells, w, scores)
        return scores

    def _score(self, cells, w, scores):
        if self.exact:
            score_exact(self, cells, w, scores)
        else:
            _lib.call("acav_mi_dense_score", self._engine, _lib.ptr(cells), w, _lib.ptr(scores), None,
                      _lib.stream_ptr(self.device))

    def calc_measure(self):
        """mi.py:108-114."""
        scores = self.score_candidates().cpu()
        score, idx = scores.max(dim=0)                             # first maximum (CPU torch, as in the reference)
        idx = int(idx)
        candidate_idx = int(self.candidate_ids[idx])
        self._add_cells(self._cand_cells[idx:idx + 1].contiguous())                      # update_cache :100-102
        self.remove_idx_all(idx)
        return float(score), candidate_idx

    def remove_idx_all(self, idx):
        """mi.py:104-125 -- order-preserving removal."""
        self.candidate_ids = torch.cat((self.candidate_ids[:idx], self.candidate_ids[idx + 1:]), dim=0)
        self._cand_cells = torch.cat((self._cand_cells[:idx], self._cand_cells[idx + 1:]), dim=0).contiguous()

    def run_greedy(self, subset_size, start_indices, intermediate_target=None, verbose=False, log_every=1,
                   log_times=None, node_rank=None, pid=None):
        """mi.py:150-192.  `start_indices` seed S but are NOT counted into the table, and the loop runs
        ``range(len(start_indices), subset_size - 1)`` -- both kept as in the reference."""
        S, GAIN, LOOKUPS, timelapse = start_indices, [], [], []
        greedy_start_time = time.time()
        for _ in range(len(start_indices), subset_size - 1):
            start_time = time.time()
            score, idx = self.calc_measure()
            timelapse.append(time.time() - start_time)
            S.append(idx)
            GAIN.append(score)
            LOOKUPS.append(0)
        if verbose:
            print("Time Consumed: {} seconds".format(time.time() - greedy_start_time))
        return (S, GAIN, timelapse, LOOKUPS)


class EfficientAMI(EfficientMI):
    """The reference's adjusted-MI measure ``ami`` (``measures/mi.py:212-262``): same greedy loop, every remaining
    candidate scored with ``(MI - EMI) / max(generalized_mean(H_a, H_b) - EMI, eps)`` of (table + candidate), EMI being
    the reference's one-term-per-cell expression (:217-231).  One ``acav_mi_dense_score_ami`` call per iteration:
    O(P*C*C) row / column sums of the EMI terms, then O(1) per candidate and pair, in fp64 (the reference spends
    O(W*P*C*C) fp32 lgamma evaluations whose differences cancel to ~n*log(n)*2^-24; agreement is 1e-6 relative at
    the sizes the reference can run, ``tests/test_zz_ami_gpu.py``)."""

    _AVERAGE = {'arithmetic': 0, 'max': 1, 'min': 2}

    def _score(self, cells, w, scores):
        method = self._AVERAGE.get(self.average_method, 0)                      # generalized_mean mi.py:200-209
        _lib.call("acav_mi_dense_score_ami", self._engine, _lib.ptr(cells), w, method, _lib.ptr(scores), None,
                  _lib.stream_ptr(self.device))
