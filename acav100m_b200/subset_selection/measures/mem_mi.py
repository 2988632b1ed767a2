"""``EfficientMemMI`` -- exact greedy mutual-information selection on a B200.

Mirror of the reference's ``EfficientMemMI`` (``subset_selection/code/measures/mi.py:284-412``) with
the greedy loop of ``EfficientMI.run_greedy`` (:150-192): same constructor keywords, ``init(pairs,
candidates)``, ``add_samples(ids)``, ``run_greedy(...) -> (S, GAIN, timelapse, LOOKUPS)``, same
selected indices and the same fp32 scores bit for bit.  All per-iteration work (score every
remaining candidate, first arg-max, table update, removal) happens inside ``libacav_b200.so``.

One clustering pair (P = 1, the K_a x K_v audio-visual table of BASELINE.json) runs on the persistent
kernels of ``acav_mi_*``; several pairs (the reference's `combination` / `bipartite` / `diagonal` pairings,
P = 45 by default) run on ``acav_mi_pairs_*`` (``pairs_engine.py``, csrc/mi_pairs.cu), the mean over pairs
added in torch's CPU summation order so that scores stay bit-identical.
New (not in the reference): ``shard=(rank, world)`` splits the candidate list into contiguous ranges
across ranks with one small all-gather per iteration; every rank returns the same S and GAIN.
"""
import contextlib
import time

import numpy as np
import torch

from ... import _lib, parallel
from . import tables
from .pairs_engine import PairsEngine


class EfficientMemMI:
    _LOOP_NAMES = {_lib.MI_LOOP_KERNELS: "kernels", _lib.MI_LOOP_PERSISTENT: "persistent",
                   _lib.MI_LOOP_CELLS: "cells", _lib.MI_LOOP_BYTES: "bytes"}

    def __init__(self, assignments, measure_type='mutual_info', average_method='arithmetic',
                 ncentroids=20, device=None, shard=None, loop='auto', **kwargs):
        self.average_method = average_method.lower()
        self.ncentroids = int(ncentroids)
        if torch.is_tensor(assignments):
            self.assignments = assignments.to(torch.long)
        else:
            self.assignments = torch.from_numpy(np.asarray(assignments)).to(torch.long)   # V x D (:24)
        self.eps = tables.EPS
        dev = device if device not in (None, 'cpu', 'cuda') else None
        self.device = _lib.require_cuda(dev)
        self.shard = shard
        self.loop = loop
        self._engine = None
        self._pairs = None                        # PairsEngine when P > 1
        self._picked = 0

    # -- setup -----------------------------------------------------------------------------------

    def init(self, clustering_combinations, candidates):
        """reference :27-30 -- empty table + candidate list."""
        self.combinations = [tuple(p) for p in clustering_combinations]
        if not self.combinations or any(len(p) != 2 for p in self.combinations):
            raise ValueError("every clustering combination must name two columns, got %r" % (self.combinations,))
        self.init_candidates(candidates)
        self.init_cache()

    def init_candidates(self, candidates):
        """``calc_N`` :285-295 -- (c1, c2) of every candidate, in list order."""
        if torch.is_tensor(candidates):
            self.candidate_ids = candidates.to(torch.long).cpu()
        else:
            self.candidate_ids = torch.as_tensor(np.asarray(candidates, dtype=np.int64))
        W = self.candidate_ids.numel()
        lo, hi = 0, W
        self._dist = None
        if self.shard is not None:
            import torch.distributed as dist
            rank, world = self.shard
            lo, hi = parallel.shard_bounds(W, rank, world)
            self._dist = dist
        self._range = (lo, hi)
        self._W = W
        ids = self.candidate_ids[lo:hi]
        a = self.assignments
        if len(self.combinations) > 1:            # rows of all clustering ids; PairsEngine keeps the columns it needs
            self._cells = None
            self._rows = a.index_select(0, ids.to(a.device))
            return
        pair = self.combinations[0]
        if a.device.type == 'cuda':
            cells = a.index_select(0, ids.to(a.device))[:, list(pair)].to(self.device).contiguous()
        else:
            cells = a.index_select(0, ids)[:, list(pair)].contiguous().to(self.device)
        if cells.numel() and (int(cells.min()) < 0 or int(cells.max()) >= self.ncentroids):
            raise ValueError("cluster ids must lie in [0, ncentroids)")
        self._cells = cells

    def init_from_cells(self, clustering_combinations, cells, w_global=None, lo=0, max_picks=None, id_offset=0,
                        all_columns=False):
        """Fast setup for big lists, skipping the per-candidate python objects of ``init`` (run_greedy.py:33 builds
        ``list(range(V))``).  One pair: `cells` is this rank's int64 [w, 2] tensor of (c1, c2) in list order -- or, with
        `all_columns`, the [w, D] tensor of all clustering ids, from which the pair's two columns are taken.  Several
        pairs: `cells` is the int64 [w, D] tensor of ALL clustering ids per candidate (the columns the pairs index).
        Pinned host or device memory; covers positions [lo, lo + w) of a candidate list of `w_global` entries whose
        clip ids are ``position + id_offset``."""
        self.combinations = [tuple(p) for p in clustering_combinations]
        if not self.combinations or any(len(p) != 2 for p in self.combinations):
            raise ValueError("every clustering combination must name two columns, got %r" % (self.combinations,))
        w = cells.shape[0]
        self._W = int(w_global) if w_global is not None else w
        self._range = (int(lo), int(lo) + w)
        self.candidate_ids = None
        self._id_offset = int(id_offset)
        self._dist = None
        if self.shard is not None:
            import torch.distributed as dist
            self._dist = dist
        if len(self.combinations) > 1:
            self._cells, self._rows = None, cells
        else:
            if all_columns:
                cells = cells[:, list(self.combinations[0])]
            self._cells = cells.to(self.device, non_blocking=True).contiguous()
        self.init_cache(max_picks=max_picks if max_picks is not None else min(self._W + 8, (1 << 24) - 8))

    def launches_per_iteration(self):
        if self._pairs is not None:
            return self._pairs.launches_per_iteration()
        if self._dist is not None and not self._nvlink:
            return 4                                   # gain, scan, emit, apply (+ one NCCL all-gather)
        return 3 if self._loop_mode() == _lib.MI_LOOP_KERNELS else 0      # persistent / cells: one launch per select()

    def loop_name(self):
        if self._pairs is not None:
            return "pairs" + ("+allgather" if self._dist is not None else "")
        if self._dist is not None:
            if self._nvlink:
                return self._LOOP_NAMES[self._loop_mode()] + "+nvlink-mailbox"
            return "kernels+allgather"
        return self._LOOP_NAMES[self._loop_mode()]

    def init_cache(self, max_picks=None):
        """``init_cache`` :32-39, :297-308 on the device."""
        self._release()
        lo, hi = self._range
        C = self.ncentroids
        if max_picks is None:
            max_picks = min(self._W + 8, (1 << 24) - 8)
        self._max_picks = int(max_picks)
        self._picked = 0
        self._nvlink = False
        self._ready_mode = None
        if len(self.combinations) > 1:
            world = self.shard[1] if self.shard is not None else 1
            self._pairs = PairsEngine(self.device, C, self.combinations, self._rows, lo, self._W, self._max_picks,
                                      dist=self._dist, world=world)
            return
        handle = _lib.c_vp()
        with torch.cuda.device(self.device):
            st = _lib.stream_ptr(self.device)
            _lib.call("acav_mi_create", _lib.ctypes.byref(handle), hi - lo, C, C, self._max_picks, lo)
            self._engine = handle
            _lib.call("acav_mi_load_candidates", handle, _lib.ptr(self._cells), st)
            self._logs = tables.log_table_device(self._max_picks + 4, self.device)
            consts = tables.empty_table_constants(C)
            _lib.call("acav_mi_set_tables", handle, _lib.ptr(self._logs), self._logs.numel(),
                      consts.ctypes.data_as(_lib.c_vp), st)
        if self._dist is not None and self._loop_mode() != _lib.MI_LOOP_KERNELS:
            self._connect_ranks()

    def _all_ranks_ok(self, ok):
        """True iff `ok` holds on every rank (one tiny all-reduce; ranks must agree before any of them enters a
        persistent kernel that waits for its peers)."""
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=self.device)
        self._dist.all_reduce(flag, op=self._dist.ReduceOp.MIN)
        return bool(flag.item())

    def _connect_ranks(self):
        """Exchange the mailbox IPC handles so the persistent kernel can push each iteration's winner
        straight into the peers' memory over NVLink (include/acav_b200.h, acav_mi_comm_*).  If any rank cannot
        map its peers (no CUDA IPC / no P2P: other node, MIG, container without shared /dev/shm) every rank falls
        back to the 3-kernel loop with one 16-byte all-gather per iteration."""
        rank, world = self.shard
        n = _lib.load().acav_mi_comm_handle_bytes()
        mine = (_lib.ctypes.c_ubyte * n)()
        err = None
        with torch.cuda.device(self.device):
            try:
                _lib.call("acav_mi_comm_export", self._engine, world, rank, mine)
            except _lib.AcavError as e:
                err = e
            local = torch.tensor(list(bytes(mine)), dtype=torch.uint8, device=self.device)
            gathered = torch.empty(world * n, dtype=torch.uint8, device=self.device)
            self._dist.all_gather_into_tensor(gathered, local)
            if self._all_ranks_ok(err is None):
                blob = bytes(gathered.cpu().numpy().tobytes())
                handles = (_lib.ctypes.c_ubyte * (world * n)).from_buffer_copy(blob)
                try:
                    _lib.call("acav_mi_comm_connect", self._engine, handles)
                except _lib.AcavError as e:
                    err = e
            self._nvlink = self._all_ranks_ok(err is None)
            self._nvlink_error = err
            self._dist.barrier()

    def _release(self):
        if self._pairs is not None:
            self._pairs.release()
            self._pairs = None
        if self._engine is not None:
            _lib.load().acav_mi_destroy(self._engine)
            self._engine = None

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    # -- operator ------------------------------------------------------------------------------

    def add_samples(self, ids):
        """reference :408-412 -- count samples into the table without selecting them."""
        if self._pairs is not None:
            for idx in ids:
                self._pairs.add_sample(self.assignments[int(idx)].tolist())
            self._picked += len(ids)
            return
        pair = self.combinations[0]
        with torch.cuda.device(self.device):
            st = _lib.stream_ptr(self.device)
            for idx in ids:
                row = self.assignments[int(idx)]
                _lib.call("acav_mi_add_sample", self._engine, int(row[pair[0]]), int(row[pair[1]]), st)
        self._picked += len(ids)

    def _loop_mode(self):
        if self.loop in ('kernels', _lib.MI_LOOP_KERNELS):
            return _lib.MI_LOOP_KERNELS
        if self.loop in ('persistent', _lib.MI_LOOP_PERSISTENT):
            return _lib.MI_LOOP_PERSISTENT
        if self.loop in ('cells', _lib.MI_LOOP_CELLS):
            return _lib.MI_LOOP_CELLS
        if self.loop in ('bytes', _lib.MI_LOOP_BYTES):
            return _lib.MI_LOOP_BYTES
        # "auto": the cheapest exact loop for this shape, by the measured cost per iteration on B200
        #   cell index    ~ 8.8 us * (K / 1024)^2         (scans the K_a x K_v cells)
        #   byte stream   ~ 14.5 us + 0.18 ns * W_local   (one byte per remaining candidate)
        #   2-byte stream ~ 14.5 us + 0.43 ns * W_local   (both measured 4000 picks into a run, W_local = 1.25e7 and 1e8)
        # then whatever else the shape supports; a layout that cannot be built for the concrete list (ACAV_E_UNSUPPORTED
        # from acav_mi_prepare) moves on to the next one (select()).
        return self._auto_order()[getattr(self, "_auto_skip", 0)]

    def _auto_order(self):
        C = self.ncentroids
        world = self.shard[1] if self.shard is not None else 1
        w_local = max(self._W // world, 1)
        cost = {_lib.MI_LOOP_CELLS: 8.8 * (C / 1024.0) ** 2,
                _lib.MI_LOOP_BYTES: 14.5 + 0.18e-3 * w_local,
                _lib.MI_LOOP_PERSISTENT: 14.5 + 0.43e-3 * w_local}
        lib = _lib.load()
        order = [m for m in sorted(cost, key=cost.get) if lib.acav_mi_loop_supported(C, C, m)]
        return order + [_lib.MI_LOOP_KERNELS]

    def check_status(self):
        """Raise if the last persistent launch gave up waiting for a peer GPU (synchronises the stream)."""
        if self._engine is None or not self._nvlink:
            return
        st = _lib.ctypes.c_int32(0)
        with torch.cuda.device(self.device):
            _lib.call("acav_mi_status", self._engine, _lib.ctypes.byref(st), _lib.stream_ptr(self.device))
        if st.value != 0:
            raise RuntimeError("greedy-MI persistent loop stopped: a peer GPU's winner did not arrive within the "
                               "spin limit (a rank failed or ran a different number of iterations)")

    def _result_views(self, n):
        """(positions int64[n], gains fp32[n]) carved out of a block allocated 64 K picks at a time: two allocator calls
        less in front of every launch of a persistent loop (a select(20) is ~0.7 ms of GPU time)."""
        blk = getattr(self, "_res_blk", None)
        if blk is None or blk[2] + n > blk[0].numel():
            cap = max(int(n), 65536)
            blk = [torch.empty(cap, dtype=torch.int64, device=self.device),
                   torch.empty(cap, dtype=torch.float32, device=self.device), 0]
            self._res_blk = blk
        lo = blk[2]
        blk[2] = lo + int(n)
        return blk[0][lo:lo + n], blk[1][lo:lo + n]

    def _device_guard(self):
        """torch.cuda.device(self.device), or nothing when that device is current already."""
        dev = torch.device(self.device)
        if dev.type == "cuda" and (dev.index is None or dev.index == torch.cuda.current_device()):
            return contextlib.nullcontext()
        return torch.cuda.device(dev)

    def select(self, n_picks):
        """Run `n_picks` greedy iterations; returns (positions int64[n] in the candidate list,
        gains fp32[n]) as device tensors, without a host sync."""
        if self._picked + n_picks + 2 > self._max_picks:
            raise RuntimeError("engine was sized for %d picks" % self._max_picks)
        if self._pairs is not None:
            self._picked += n_picks
            return self._pairs.select(n_picks)
        pos, gain = self._result_views(n_picks)
        with self._device_guard():
            st = _lib.stream_ptr(self.device)
            if self._dist is None or self._nvlink:
                if self._nvlink and getattr(self, "_ready_mode", None) != self._loop_mode():
                    # every rank must have its layout ready (and have got here) before any enters the kernel
                    try:
                        _lib.call("acav_mi_prepare", self._engine, self._loop_mode(), st)
                        err = None
                    except _lib.AcavError as e:
                        err = e
                    while not self._all_ranks_ok(err is None):
                        # "auto": every rank moves on to the next loop together; a named loop fails loudly
                        if self.loop != 'auto' or self._loop_mode() == _lib.MI_LOOP_KERNELS:
                            raise RuntimeError("greedy-MI setup failed on %s rank: %s"
                                               % ("this" if err else "another", err))
                        self._auto_skip = getattr(self, "_auto_skip", 0) + 1
                        try:
                            _lib.call("acav_mi_prepare", self._engine, self._loop_mode(), st)
                            err = None
                        except _lib.AcavError as e:
                            err = e
                    self._ready_mode = self._loop_mode()
                if self._dist is None and self.loop == 'auto' and getattr(self, "_ready_mode", None) is None:
                    # one GPU: build the layout of the chosen loop now; if this list does not fit it, take the next loop
                    while True:
                        try:
                            _lib.call("acav_mi_prepare", self._engine, self._loop_mode(), st)
                            break
                        except _lib.AcavError as e:
                            if e.status != _lib.E_UNSUPPORTED or self._loop_mode() == _lib.MI_LOOP_KERNELS:
                                raise
                            self._auto_skip = getattr(self, "_auto_skip", 0) + 1
                    self._ready_mode = self._loop_mode()
                _lib.call("acav_mi_run", self._engine, n_picks, _lib.ptr(pos), _lib.ptr(gain),
                          self._loop_mode(), st)
            else:
                dev, eng = self.device, self._engine

                class _Engine:
                    def local_best(self, out_pair):
                        _lib.call("acav_mi_local_best", eng, _lib.ptr(out_pair), st)

                    def apply(self, all_pairs, world, i):
                        _lib.call("acav_mi_apply", eng, _lib.ptr(all_pairs), world,
                                  _lib.c_vp(pos.data_ptr() + 8 * i), _lib.c_vp(gain.data_ptr() + 4 * i), st)

                parallel.sharded_greedy(
                    _Engine(), self._dist, self.shard[1], n_picks,
                    lambda: torch.empty(2, dtype=torch.int64, device=dev),
                    lambda world: torch.empty(2 * world, dtype=torch.int64, device=dev))
        self._picked += n_picks
        return pos, gain

    def run_greedy(self, subset_size, start_indices, intermediate_target=None,
                   verbose=False, log_every=1, log_times=None,
                   node_rank=None, pid=None):
        """reference :150-192.  `start_indices` seed S but are NOT counted into the table, and the
        loop runs ``range(len(start_indices), subset_size - 1)`` -- both kept as in the reference."""
        S = start_indices
        n_picks = max(subset_size - 1 - len(start_indices), 0)
        if n_picks > self._W:
            raise RuntimeError("cannot pick %d of %d candidates" % (n_picks, self._W))
        t0 = time.time()
        pos, gain = self.select(n_picks)
        self.check_status()
        pos = pos.cpu()
        gains = gain.cpu().tolist()
        elapsed = time.time() - t0
        S.extend((pos + getattr(self, "_id_offset", 0) if self.candidate_ids is None else self.candidate_ids[pos]).tolist())
        GAIN = [float(g) for g in gains]
        timelapse = [elapsed / n_picks] * n_picks if n_picks else []
        LOOKUPS = [0] * n_picks
        if verbose:
            print("Time Consumed: {} seconds".format(elapsed))
        return (S, GAIN, timelapse, LOOKUPS)

    def read_state(self):
        """Table counts and running sums (N [C,C], a [C], b [C], {NlogN, aloga, blogb, n}) for tests; with P > 1
        pairs every item gains a leading P axis."""
        if self._pairs is not None:
            return self._pairs.read_state()
        C = self.ncentroids
        N = torch.empty(C * C, dtype=torch.int32, device=self.device)
        a = torch.empty(C, dtype=torch.int32, device=self.device)
        b = torch.empty(C, dtype=torch.int32, device=self.device)
        sums = torch.empty(4, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.call("acav_mi_read_state", self._engine, _lib.ptr(N), _lib.ptr(a), _lib.ptr(b),
                      _lib.ptr(sums), _lib.stream_ptr(self.device))
        return N.view(C, C).cpu(), a.cpu(), b.cpu(), sums.cpu()
