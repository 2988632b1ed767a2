"""``EfficientBatchMI`` -- the reference CLI's default measure (stochastic batch greedy) on a B200.

Mirror of ``subset_selection/code/measures/batch.py:10-260``: same constructor keywords, ``init``,
``add_samples``, ``run_greedy -> (S, GAIN, timelapse, LOOKUPS)``.  Per iteration the reference reshuffles
ALL remaining candidates with ``torch.randperm`` on the CPU generator, scores the first B = 20 with the
dense MI of (table + one-hot), keeps the top k = 4 and re-appends the losers in ascending id order.

What runs where: the candidate bookkeeping (randperm, slicing, ``unique``) stays on torch CPU tensors
exactly as in the reference -- the shuffle must consume the same generator stream to pick the same
batches, and it, not the scoring, sets the pace (38 ms per iteration at 1 M candidates, SURVEY section
3.5).  The scoring and the table update run in ``libacav_b200.so`` (``acav_mi_dense_*``).

Parity: which of several near-equal candidates ``topk`` returns in the reference is decided by the fp32 rounding of
its dense sum, so with ``exact=True`` (default) the B candidates are scored by ``acav_mi_dense_score_exact``: the
reference's own dense evaluation, five fp32 roundings per cell, summed in the order of torch's CPU reduction kernel
and averaged over the pairs in that kernel's order -- the same bits, hence the same top-k and, with the same seed,
the same S as the reference's CPU run (``tests/test_batch_mi_gpu.py`` asserts it on goldens written by the
unmodified reference).  ``exact=False`` scores in O(B*P) from fp64 running sums (~1e-6 relative agreement, picks
identical while no near-tie decides one).
"""
import math
import time

import numpy as np
import torch

from ... import _lib
from . import tables
from .dense_mi import score_exact


class EfficientBatchMI:
    def __init__(self, assignments, measure_type='mutual_info', average_method='arithmetic',
                 ncentroids=20, batch_size=1, selection_size=1, device='cpu', keep_unselected=False, exact=True,
                 **kwargs):
        self.average_method = average_method.lower()
        self.ncentroids = int(ncentroids)
        self.assignments = torch.from_numpy(np.asarray(assignments)).to(torch.long)      # V x D (mi.py:24)
        self.eps = tables.EPS
        self.B = batch_size
        self.k = selection_size
        self.keep_unselected = keep_unselected
        self.device = _lib.require_cuda(device if device not in (None, 'cpu', 'cuda') else None)
        self.exact = exact
        self._engine = None
        self._n_added = 0

    def init(self, clustering_combinations, candidates):
        """mi.py:27-30 with batch.py:20-27."""
        self._n_added = 0
        self.combinations = [tuple(p) for p in clustering_combinations]
        self._pair_ids = torch.as_tensor(self.combinations, dtype=torch.long)            # P x 2
        self.candidate_ids = torch.as_tensor(np.asarray(candidates, dtype=np.int64))     # batch.py:20-22
        self._release()
        handle = _lib.c_vp()
        with torch.cuda.device(self.device):
            _lib.call("acav_mi_dense_create", _lib.ctypes.byref(handle), len(self.combinations), self.ncentroids,
                      _lib.stream_ptr(self.device))
        self._engine = handle

    def _release(self):
        if self._engine is not None:
            _lib.load().acav_mi_dense_destroy(self._engine)
            self._engine = None

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    # -- device calls ------------------------------------------------------------------------------

    def _cells(self, ids):
        """(c1, c2) per clustering pair of the given clips -> int64 [m, P, 2] on the device."""
        rows = self.assignments.index_select(0, ids)                                     # get_assignments mi.py:41-45
        cells = rows[:, self._pair_ids]                                                  # [m, P, 2]
        if cells.numel() and (int(cells.min()) < 0 or int(cells.max()) >= self.ncentroids):
            raise ValueError("cluster ids must lie in [0, ncentroids)")
        return cells.contiguous().to(self.device)

    def add_samples(self, ids):
        """batch.py:190-193 -- count clips into the tables."""
        ids = torch.as_tensor(ids, dtype=torch.long)
        cells = self._cells(ids)
        self._n_added += int(cells.shape[0])
        with torch.cuda.device(self.device):
            _lib.call("acav_mi_dense_add", self._engine, _lib.ptr(cells), cells.shape[0], _lib.stream_ptr(self.device))

    def score_batch(self, batch_ids):
        """``operate_block`` + ``mean(dim=-1)`` (batch.py:123-130,144): fp32 score per clip, on the host
        (the reference also brings them back, batch.py:135)."""
        cells = self._cells(batch_ids)
        scores = torch.empty(cells.shape[0], dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            if self.exact:
                score_exact(self, cells, cells.shape[0], scores)
            else:
                _lib.call("acav_mi_dense_score", self._engine, _lib.ptr(cells), cells.shape[0], _lib.ptr(scores), None,
                          _lib.stream_ptr(self.device))
        return scores.cpu()

    # -- host bookkeeping, as in the reference -----------------------------------------------------

    def shuffle_candidate_ids(self):
        """batch.py:29-32."""
        idx = torch.randperm(self.candidate_ids.shape[0])
        self.candidate_ids = self.candidate_ids.index_select(0, idx)

    def calc_ids(self, scores):
        """batch.py:143-150 (scores are already averaged over the pairs)."""
        k = self.k
        if scores.shape[0] < self.B:
            k = math.floor(self.B / self.k * scores.shape[0])
        return scores.topk(k=k, dim=0)

    def get_unselected(self, orig, selected):
        """batch.py:167-171."""
        uniques, counts = torch.cat((orig, selected), dim=0).unique(return_counts=True)
        return uniques[counts == 1]

    def update_candidates(self, chosen):
        """batch.py:156-165."""
        batch = self.candidate_ids[:self.B]
        self.candidate_ids = self.candidate_ids[self.B:]
        if self.keep_unselected:
            unselected = self.get_unselected(batch, chosen)
            assert unselected.shape[0] + chosen.shape[0] == batch.shape[0], \
                'wrong unselected_size: unselected {} + {} != {}'.format(unselected.shape[0], chosen.shape[0],
                                                                        batch.shape[0])
            self.candidate_ids = torch.cat((self.candidate_ids, unselected), dim=0)

    def modify_k(self, subset_size):
        """batch.py:173-188."""
        term = self.B * subset_size / self.assignments.shape[0]
        k = self.k
        if k < term and not self.keep_unselected:
            print("k={} is too small to get {} samples from {} datapoints with batch_size {}".format(
                k, subset_size, self.assignments.shape[0], self.B))
            k = math.ceil(term)
            print("resizing k to {}".format(k))
        return k

    def calc_measure_batch(self):
        """batch.py:132-137 / block_operate :93-121."""
        self.shuffle_candidate_ids()
        batch = self.candidate_ids[:self.B]
        scores, ids = self.calc_ids(self.score_batch(batch))
        chosen = batch.index_select(0, ids)
        self.add_samples(chosen)                                   # update_cache :152-154
        self.update_candidates(chosen)
        return scores, chosen, 1

    def run_greedy(self, subset_size, start_indices, intermediate_target=None, verbose=False, log_every=1,
                   log_times=None, node_rank=None, pid=None):
        """batch.py:195-260.  The start clips are counted into the tables but never into S."""
        S, GAIN, LOOKUPS, timelapse = [], [], [], []
        self.k = self.modify_k(subset_size)
        self.add_samples(start_indices)
        greedy_start_time = time.time()
        dataset_size = self.candidate_ids.shape[0]
        while len(S) < subset_size:
            start_time = time.time()
            scores, chosen, lookup = self.calc_measure_batch()
            timelapse.append(time.time() - start_time)
            S += chosen.tolist()
            GAIN += scores.tolist()
            LOOKUPS.append(lookup)
            if self.keep_unselected:
                assert self.candidate_ids.shape[0] + len(S) == dataset_size, \
                    "dataset size mismatch: {} + {} != {}".format(self.candidate_ids.shape[0], len(S), dataset_size)
        S = S[:subset_size]
        if verbose:
            print("Time Consumed: {} seconds".format(time.time() - greedy_start_time))
        return (S, GAIN, timelapse, LOOKUPS)
