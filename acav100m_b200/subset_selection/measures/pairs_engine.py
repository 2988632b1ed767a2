"""Device engine behind ``EfficientMemMI`` for P > 1 clustering pairs (``acav_mi_pairs_*``, csrc/mi_pairs.cu).

The reference keeps, per candidate, the P cell coordinates ``[W, P, 2]`` int64 (``calc_N``, measures/mi.py:285-295) and
per iteration gathers ``[P, W, C]`` floats (``small_gather`` :360-366).  Here a candidate is the row of its D' distinct
cluster ids (uint16 columns on the device, D' = number of clustering columns the pairs mention) and an iteration is a
P*C*C score table plus one streaming pass -- see the kernel file.  Scores (the mean over pairs in torch's CPU summation
order) and picks are bit-identical to the reference's.
"""
import numpy as np
import torch

from ... import _lib, parallel
from . import tables


class PairsEngine:
    def __init__(self, device, ncentroids, pairs, rows, lo, w_global, max_picks, dist=None, world=1):
        """rows: int64 [w, D] cluster ids of this rank's candidates in list order (host or device tensor, all D
        clustering columns); pairs: [(col1, col2)] over those D columns; lo: global position of rows[0]."""
        self.device = device
        self.C = int(ncentroids)
        self.pairs = [tuple(int(c) for c in p) for p in pairs]
        self.P = len(self.pairs)
        used = sorted({c for p in self.pairs for c in p})
        self.columns = used                                          # clustering column of every engine column
        remap = {c: i for i, c in enumerate(used)}
        self.engine_pairs = np.ascontiguousarray([[remap[a], remap[b]] for a, b in self.pairs], dtype=np.int32)
        self.D = len(used)
        self.max_picks = int(max_picks)
        self._dist, self._world = dist, int(world)
        ids = rows[:, used] if rows.shape[1] != len(used) or used != list(range(rows.shape[1])) else rows
        ids = ids.to(device).contiguous()
        if ids.numel() and (int(ids.min()) < 0 or int(ids.max()) >= self.C):
            raise ValueError("cluster ids must lie in [0, ncentroids)")
        self.w = ids.shape[0]
        handle = _lib.c_vp()
        self._h = None
        with torch.cuda.device(device):
            st = _lib.stream_ptr(device)
            _lib.call("acav_mi_pairs_create", _lib.ctypes.byref(handle), self.w, self.D, self.C, self.P,
                      self.engine_pairs.ctypes.data_as(_lib.c_vp), self.max_picks, int(lo))
            self._h = handle
            try:
                _lib.call("acav_mi_pairs_load_candidates", handle, _lib.ptr(ids), st)
                self._logs = tables.log_table_device(self.max_picks + 4, device)
                consts = np.ascontiguousarray(tables.pair_table_constants(self.P, self.C))
                _lib.call("acav_mi_pairs_set_tables", handle, _lib.ptr(self._logs), self._logs.numel(),
                          consts.ctypes.data_as(_lib.c_vp), st)
                torch.cuda.current_stream(device).synchronize()      # `ids` may be freed once packed
            except Exception:
                self.release()
                raise
        self.record_words = int(_lib.load().acav_mi_pairs_record_words(handle))

    def release(self):
        if getattr(self, "_h", None) is not None:
            _lib.load().acav_mi_pairs_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass

    def add_sample(self, row):
        """row: the sample's D cluster ids (all clustering columns)."""
        ids = np.ascontiguousarray([int(row[c]) for c in self.columns], dtype=np.int64)
        with torch.cuda.device(self.device):
            _lib.call("acav_mi_pairs_add_sample", self._h, ids.ctypes.data_as(_lib.c_vp), _lib.stream_ptr(self.device))

    def launches_per_iteration(self):
        return 4                                                     # gain table, scan, emit, apply

    def select(self, n_picks):
        pos = torch.empty(n_picks, dtype=torch.int64, device=self.device)
        gain = torch.empty(n_picks, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            st = _lib.stream_ptr(self.device)
            if self._dist is None:
                _lib.call("acav_mi_pairs_run", self._h, n_picks, _lib.ptr(pos), _lib.ptr(gain), st)
            else:
                dev, h, words = self.device, self._h, self.record_words

                class _Engine:
                    def local_best(self, out_rec):
                        _lib.call("acav_mi_pairs_local_best", h, _lib.ptr(out_rec), st)

                    def apply(self, all_recs, world, i):
                        _lib.call("acav_mi_pairs_apply", h, _lib.ptr(all_recs), world,
                                  _lib.c_vp(pos.data_ptr() + 8 * i), _lib.c_vp(gain.data_ptr() + 4 * i), st)

                parallel.sharded_greedy(
                    _Engine(), self._dist, self._world, n_picks,
                    lambda: torch.empty(words, dtype=torch.int64, device=dev),
                    lambda world: torch.empty(words * world, dtype=torch.int64, device=dev))
        return pos, gain

    def read_state(self):
        P, C = self.P, self.C
        N = torch.empty(P * C * C, dtype=torch.int32, device=self.device)
        a = torch.empty(P * C, dtype=torch.int32, device=self.device)
        b = torch.empty(P * C, dtype=torch.int32, device=self.device)
        sums = torch.empty(P * 4, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.call("acav_mi_pairs_read_state", self._h, _lib.ptr(N), _lib.ptr(a), _lib.ptr(b), _lib.ptr(sums),
                      _lib.stream_ptr(self.device))
        return N.view(P, C, C).cpu(), a.view(P, C).cpu(), b.view(P, C).cpu(), sums.view(P, 4).cpu()
