"""``KMeans`` -- mini-batch SGD k-means with the reference's operator API, computed on a B200.

Mirror of ``clustering/code/sgd_clustering.py:10-129`` of the reference: same constructor, same
attributes (``centers``, ``counts``, ``count``, ``lr``, ``initial_rounds``, ``reinit``, ``fallback``),
same methods (``to``, ``initialize``, ``add``, ``calc_best``, ``get_attrs``, ``load``), same
multi-GPU semantics (replicated centers, per-rank batch slice, sum of histograms and deltas).
Every tensor operation of the reference is replaced by a call into ``libacav_b200.so``; there is
no torch math and no CPU fallback on the compute path.

Differences a user can observe (all documented in DESIGN.md):
  * the per-step ``all_gather`` of the batch (reference :97, used only for ``len``) is dropped;
  * ``fallback`` is kept on the device (no host sync per step, reference :116-119) and read lazily;
  * ``add`` / ``calc_best`` take ``sync=False`` to return the mean distance as a 0-dim device tensor
    instead of a python float (the reference's ``.item()`` at :79 stalls the stream every call);
  * past the warm-up ``add`` replays the whole step (about 16 kernels, plus the two all-reduces on several GPUs)
    from a CUDA graph cached per batch address (``graph='auto'``): same kernels, same arguments, same results
    bit for bit -- one launch instead of ~20 host calls.  ``last_best`` then aliases a buffer that the next step
    overwrites.
"""
import collections

import numpy as np
import torch

from .. import _lib, parallel

_GRAPH_CACHE = 512          # captured steps kept per KMeans object (one per distinct batch address)
_THR_RING = 64              # pinned host slots for the per-step threshold (H2D copy ahead of every replay)


class KMeans:
    def __init__(self, args=None, d=None, k=None, lr=1e-2,
                 initial_rounds=10, reinit=(.7, 5.0), saved_dt=None,
                 assign_mode="auto", warmup_rng="cpu", tile_variant=0, graph="auto", comm="auto"):
        self._ws = None
        self.comm = comm                              # multi-GPU exchange: 'auto' / 'p2p' (NVLink peer memory) or 'nccl'
        self._comm = None                             # acav_kmeans_comm_t once connected; False = use NCCL
        self.graph = graph                            # 'auto' / True: CUDA-graph replay of the steady-state step; False: eager
        self._gs = None
        self.tile_variant = tile_variant              # _lib.TILE_*: which tcgen05 distance-GEMM kernel (0 = by shape)
        self._ws_batch = 0
        self._fallback_base = 0
        self._fallback_dev = None
        self.assign_mode = assign_mode
        self.warmup_rng = warmup_rng
        if saved_dt is not None:
            self.load_from_saves(saved_dt)
        else:
            self.args = args
            self.centers = torch.rand(k, d) * 1e-5          # reference :24 (global CPU generator)
            self.counts = torch.zeros(k)
            self.count = 0
            self.lr = lr
            self.initial_rounds = initial_rounds
            self.reinit = reinit
            self.fallback = 0
            self.sequential = False

    # -- state ---------------------------------------------------------------------------------

    @property
    def fallback(self):
        """Number of lr fallbacks taken (reference :119); the device counter is read on access."""
        extra = int(self._fallback_dev.item()) if self._fallback_dev is not None else 0
        return self._fallback_base + extra

    @fallback.setter
    def fallback(self, value):
        self._fallback_base = int(value)
        if self._fallback_dev is not None:
            self._fallback_dev.zero_()

    def get_attrs(self):
        """reference :34-46."""
        self.check_status()
        return {
            'args': self.args, 'count': self.count, 'lr': self.lr,
            'initial_rounds': self.initial_rounds, 'reinit': self.reinit,
            'fallback': self.fallback, 'sequential': self.sequential,
            'centers': self.centers.cpu().numpy(), 'counts': self.counts.cpu().numpy(),
        }

    def load_from_saves(self, dt):
        """reference :48-52."""
        for key, val in dt.items():
            setattr(self, key, val)
        self.centers = torch.from_numpy(np.ascontiguousarray(self.centers, dtype=np.float32))
        self.counts = torch.from_numpy(np.ascontiguousarray(self.counts, dtype=np.float32))

    @classmethod
    def load(cls, dt):
        return cls(saved_dt=dt)

    def to(self, device):
        """reference :59-61."""
        self.centers = self.centers.to(device).contiguous()
        self.counts = self.counts.to(device).contiguous()
        self._release()

    def __getstate__(self):
        st = dict(self.__dict__)
        st['_fallback_base'] = self.fallback
        st['_fallback_dev'] = None
        st['_gs'] = None
        st['_comm'] = None
        st['_ws'] = None
        st['_ws_pair'] = None
        st['_ws_batch'] = 0
        st['centers'] = self.centers.cpu()
        st['counts'] = self.counts.cpu()
        return st

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    # -- workspace -----------------------------------------------------------------------------

    def _release(self):
        if getattr(self, "_comm", None):
            torch.cuda.synchronize(self.centers.device)
            _lib.load().acav_kmeans_comm_destroy(self._comm)
        self._comm = None
        self._release_workspace()

    def _release_workspace(self):
        """Workspace, its pair and the graphs captured over them; the peer-memory exchange (sized by k and d only)
        stays connected when a bigger batch makes the workspace grow."""
        self._gs = None
        self._release_pair()
        if self._ws is not None:
            _lib.load().acav_kmeans_destroy(self._ws)
            self._ws = None
            self._ws_batch = 0

    def _device(self):
        if self.centers.device.type != 'cuda':
            raise RuntimeError("KMeans is on %s: move it with .to('cuda') -- acav100m_b200 has no CPU path"
                               % self.centers.device)
        return self.centers.device

    def _workspace(self, b):
        dev = self._device()
        if self._ws is None or b > self._ws_batch:
            self._release_workspace()
            k, d = self.centers.shape
            cap = max(int(b), 1024)
            handle = _lib.c_vp()
            with torch.cuda.device(dev):
                _lib.call("acav_kmeans_create", _lib.ctypes.byref(handle), k, d, cap)
                if getattr(self, "tile_variant", 0):
                    _lib.call("acav_kmeans_set_tile_variant", handle, int(self.tile_variant))
            self._ws, self._ws_batch = handle, cap
        if self._fallback_dev is None or self._fallback_dev.device != dev:
            self._fallback_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        return self._ws

    def _prep_batch(self, batch):
        dev = self._device()
        batch = batch.to(device=dev, dtype=torch.float32)
        if batch.dim() != 2 or batch.shape[1] != self.centers.shape[1]:
            raise ValueError("batch must be [b, %d], got %s" % (self.centers.shape[1], tuple(batch.shape)))
        if batch.stride(1) != 1 or batch.stride(0) < batch.shape[1]:
            batch = batch.contiguous()
        return batch

    def _mode(self):
        if self.assign_mode in ("exact", _lib.ASSIGN_EXACT):
            return _lib.ASSIGN_EXACT
        if self.assign_mode in ("tensor", _lib.ASSIGN_TENSOR):
            return _lib.ASSIGN_TENSOR
        return _lib.ASSIGN_TENSOR                  # "auto": tcgen05 screen + exact re-check of near-ties

    def mode_name(self):
        return "exact" if self._mode() == _lib.ASSIGN_EXACT else "tensor"

    def launches_per_step(self):
        """CUDA kernels of this library launched by one add() past warm-up (bench.py gpu_launches)."""
        # tensor mode: centroid prep (2) + row prep + distance GEMM + classify + candidate re-check + exact kernel
        # + exact distance of the winner + mean; then the stable partition (3 kernels up to 32768 rows, else 4) and the
        # update:
        #   one GPU: 2 update kernels (they take the lr decision themselves; beyond 32768 rows + effective lr);
        #   NCCL: effective lr + 2 update kernels + apply; peer memory: histogram exchange + 2 update kernels + signal,
        #   reduce/broadcast, signal, gather
        assign = 4 if self._mode() == _lib.ASSIGN_EXACT else 9
        _, world = self._world()
        mid = 0 < self._ws_batch <= 32768
        tail = (2 if mid else 3) if world == 1 else (7 if self._comm else 4)
        part = 3 if mid else 4
        return assign + part + tail + (1 if self._gs is not None else 0)

    # -- operator ------------------------------------------------------------------------------

    @property
    def in_warmup(self):
        """reference :67."""
        return self.count < self.initial_rounds * self.centers.shape[0]

    def underused_threshold(self):
        """Right-hand side of ``counts < (count / k) ** p`` (reference :77) as torch compares it:
        the python scalar is cast to fp32 before the comparison."""
        k = self.centers.shape[0]
        return float(np.float32((self.count / k) ** self.reinit[0]))

    def _assign(self, batch, want_mean):
        """Device part of calc_best: returns (best int64[b], mean 0-dim tensor or None)."""
        dev = self._device()
        k, d = self.centers.shape
        b = batch.shape[0]
        ws = self._workspace(b)
        best = torch.empty(b, dtype=torch.int64, device=dev)
        mean = torch.empty(1, dtype=torch.float32, device=dev) if want_mean else None
        with torch.cuda.device(dev):
            st = _lib.stream_ptr(dev)
            if self.in_warmup:
                if self.warmup_rng == "cpu":
                    noise = torch.rand(k, b).to(dev, non_blocking=False)        # reference :68
                else:
                    noise = torch.rand(k, b, device=dev)
                mind = torch.empty(b, dtype=torch.float32, device=dev) if want_mean else None
                _lib.call("acav_kmeans_assign_noise", _lib.ptr(noise), k, b, _lib.ptr(best),
                          _lib.ptr(mind), _lib.ptr(mean), st)
            else:
                _lib.call("acav_kmeans_assign", ws, _lib.ptr(batch, row_strided=True), b, batch.stride(0),
                          _lib.ptr(self.centers), _lib.ptr(self.counts),
                          self.underused_threshold(), float(self.reinit[1]),
                          _lib.ptr(best), None, _lib.ptr(mean), None, self._mode(), st)
        return best, mean

    def calc_best(self, batch, sync=True, distance=True):
        """reference :63-79 -> (best LongTensor[b] on the device, mean min-distance).
        `distance=False` skips the exact-distance pass over the batch and returns None for the mean
        (the reference's callers that only label clips discard it, process_batch.py:43-48)."""
        batch = self._prep_batch(batch)
        best, mean = self._assign(batch, want_mean=distance)
        if not distance:
            return best, None
        return best, (mean.item() if sync else mean[0])

    def assign_all(self, x, chunk=131072):
        """Labels for a whole resident feature matrix (the assignment pass, run_clustering.py:225-229),
        walked in chunks over two workspaces and two streams (preparation of chunk i+1 / tensor-core kernel
        of chunk i).  Returns LongTensor[n] on the device."""
        x = self._prep_batch(x)
        dev = self._device()
        n = x.shape[0]
        best = torch.empty(n, dtype=torch.int64, device=dev)
        if n == 0:
            return best
        if self.in_warmup or self._mode() != _lib.ASSIGN_TENSOR:
            for lo in range(0, n, chunk):
                best[lo:lo + chunk] = self._assign(x[lo:lo + chunk], want_mean=False)[0]
            return best
        k, d = self.centers.shape
        chunk = int(min(chunk, n))
        if getattr(self, "_ws_pair", None) is None or self._ws_pair[2] < chunk:
            self._release_pair()
            hs = []
            with torch.cuda.device(dev):
                for _ in range(2):
                    h = _lib.c_vp()
                    _lib.call("acav_kmeans_create", _lib.ctypes.byref(h), k, d, chunk)
                    if getattr(self, "tile_variant", 0):
                        _lib.call("acav_kmeans_set_tile_variant", h, int(self.tile_variant))
                    hs.append(h)
            # second stream for the distance GEMM.  Measured on B200: the preparation of chunk i+1 and the GEMM
            # of chunk i do share the SMs, but both run at the board's power limit (~1 kW), so the pass takes
            # the SUM of their times either way (tools/km_overlap_probe.py)
            self._ws_pair = (hs[0], hs[1], chunk, torch.cuda.Stream(device=dev, priority=-1))
        ws, hot = self._ws_pair[:2], self._ws_pair[3]
        thr, r = self.underused_threshold(), float(self.reinit[1])
        main = torch.cuda.current_stream(dev)
        with torch.cuda.device(dev):
            mp = _lib.ctypes.c_void_p(main.cuda_stream)
            hp = _lib.ctypes.c_void_p(hot.cuda_stream)
            for h in ws:
                _lib.call("acav_kmeans_prepare_centers", h, _lib.ptr(self.centers), _lib.ptr(self.counts), thr, r, mp)
            done = [None, None]
            for i, lo in enumerate(range(0, n, chunk)):
                xb = x[lo:lo + chunk]
                h = ws[i % 2]
                if done[i % 2] is not None:
                    main.wait_event(done[i % 2])                   # workspace free again
                _lib.call("acav_kmeans_prepare_batch", h, _lib.ptr(xb, row_strided=True), xb.shape[0], xb.stride(0), mp)
                ready = torch.cuda.Event()
                ready.record(main)
                hot.wait_event(ready)
                _lib.call("acav_kmeans_assign_prepared", h, _lib.ptr(xb, row_strided=True), xb.shape[0],
                          xb.stride(0), _lib.ptr(self.centers), _lib.ptr(self.counts), thr, r,
                          _lib.c_vp(best.data_ptr() + 8 * lo), None, None, None, hp)
                done[i % 2] = torch.cuda.Event()
                done[i % 2].record(hot)
            main.wait_stream(hot)
            best.record_stream(hot)
        return best

    def _release_pair(self):
        pair = getattr(self, "_ws_pair", None)
        if pair is not None:
            torch.cuda.synchronize(self.centers.device)
            for h in pair[:2]:
                _lib.load().acav_kmeans_destroy(h)
            self._ws_pair = None

    @property
    def is_distributed(self):
        """reference :81-86."""
        return (self.args is not None and self.args.computation.device == 'cuda'
                and self.args.computation.num_gpus > 1)

    def _world(self):
        import torch.distributed as dist
        if self.is_distributed and dist.is_available() and dist.is_initialized():
            return dist, dist.get_world_size()
        return None, 1

    def initialize(self):
        """reference :88-92 -- average the independently drawn inits over ranks."""
        dist, world = self._world()
        if dist is not None and world > 1:
            for t in (self.centers, self.counts):
                dist.all_reduce(t)
                t.mul_(1.0 / world)

    # device hooks of add(): each is one or two C-ABI calls (overridden by the CPU protocol tests)

    def _histogram(self, batch, best):
        k = self.centers.shape[0]
        b = batch.shape[0]
        counts_b = torch.empty(k, dtype=torch.float32, device=self.centers.device)
        with torch.cuda.device(self.centers.device):
            _lib.call("acav_kmeans_histogram", self._workspace(b), _lib.ptr(best), b, _lib.ptr(counts_b),
                      _lib.stream_ptr(self.centers.device))
        return counts_b

    def _update_fused(self, batch, counts_b, lr):
        b = batch.shape[0]
        with torch.cuda.device(self.centers.device):
            _lib.call("acav_kmeans_update_fused", self._workspace(b), _lib.ptr(batch, row_strided=True), b,
                      batch.stride(0), _lib.ptr(counts_b), lr, _lib.ptr(self.centers), _lib.ptr(self.counts),
                      _lib.ptr(self._fallback_dev), _lib.stream_ptr(self.centers.device))

    def _update_local(self, batch, counts_b_global, lr, deltas=None):
        k, d = self.centers.shape
        b = batch.shape[0]
        if deltas is None:
            deltas = torch.empty(k, d, dtype=torch.float32, device=self.centers.device)
        with torch.cuda.device(self.centers.device):
            _lib.call("acav_kmeans_update_local", self._workspace(b), _lib.ptr(batch, row_strided=True), b,
                      batch.stride(0), _lib.ptr(counts_b_global), lr, _lib.ptr(self.centers),
                      _lib.ptr(self.counts), _lib.ptr(deltas), _lib.ptr(self._fallback_dev),
                      _lib.stream_ptr(self.centers.device))
        return deltas

    def _apply_deltas(self, deltas):
        with torch.cuda.device(self.centers.device):
            _lib.call("acav_kmeans_apply_deltas", _lib.ptr(self.centers), _lib.ptr(deltas), deltas.numel(),
                      _lib.stream_ptr(self.centers.device))

    def add(self, batch, sync=True, distance=True):
        """reference :94-129 (fast parallel update) -> mean min-distance of the batch
        (None with `distance=False`: the training driver discards it, run_clustering.py:171-175)."""
        batch = self._prep_batch(batch)
        dev = self._device()
        b = batch.shape[0]
        dist, world = self._world()
        lr = self.lr(self.count) if callable(self.lr) else self.lr
        if self.sequential:
            # reference :96-109 -- the slow branch works on the all-gathered batch, every rank applies every row
            if world > 1:
                gbatch = torch.empty((b * world, batch.shape[1]), dtype=batch.dtype, device=dev)
                dist.all_gather_into_tensor(gbatch, batch.contiguous())
                batch = gbatch
            best, mean = self._assign(batch, want_mean=distance)
            counts_b = self._histogram(batch, best)
            with torch.cuda.device(dev):
                _lib.call("acav_kmeans_update_sequential", self._workspace(batch.shape[0]),
                          _lib.ptr(batch, row_strided=True), batch.shape[0], batch.stride(0), _lib.ptr(counts_b),
                          float(lr), _lib.ptr(self.centers), _lib.ptr(self.counts), _lib.stream_ptr(dev))
        elif self._graphable(batch):
            best, mean = self._add_graphed(batch, float(lr), distance, dist, world)
        else:
            best, mean = self._assign(batch, want_mean=distance)
            counts_b = self._histogram(batch, best)                                                 # :113
            self._update(batch, counts_b, float(lr), dist, world)                                   # :114-127
        self.count += parallel.kmeans_global_batch(b, world)                                        # :128
        self.last_best = best
        if not distance:
            return None
        return mean.item() if sync else mean[0]

    def _peer_comm(self, dist, world):
        """NVLink peer-memory exchange for the distributed step (acav_kmeans_comm_*): created on the first
        multi-GPU step; every rank agrees (one all-reduce) whether all of them could map their peers, else all
        use NCCL.  Returns the handle or None."""
        if self._comm is not None:
            return self._comm or None
        if self.comm == "nccl" or type(self)._histogram is not KMeans._histogram:
            self._comm = False
            return None
        dev = self._device()
        k, d = self.centers.shape
        handle, err = _lib.c_vp(), None
        n = _lib.load().acav_kmeans_comm_handle_bytes()
        mine = (_lib.ctypes.c_ubyte * n)()
        with torch.cuda.device(dev):
            try:
                _lib.call("acav_kmeans_comm_create", _lib.ctypes.byref(handle), k, d, world, dist.get_rank())
                _lib.call("acav_kmeans_comm_export", handle, mine)
            except _lib.AcavError as e:
                err = e

            def all_ok(ok):
                flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
                dist.all_reduce(flag, op=dist.ReduceOp.MIN)
                return bool(flag.item())

            local = torch.tensor(list(bytes(mine)), dtype=torch.uint8, device=dev)
            gathered = torch.empty(world * n, dtype=torch.uint8, device=dev)
            dist.all_gather_into_tensor(gathered, local)
            if all_ok(err is None):
                handles = (_lib.ctypes.c_ubyte * (world * n)).from_buffer_copy(gathered.cpu().numpy().tobytes())
                try:
                    _lib.call("acav_kmeans_comm_connect", handle, handles)
                except _lib.AcavError as e:
                    err = e
            ok = all_ok(err is None)
            dist.barrier()
        if not ok:
            if handle:
                _lib.load().acav_kmeans_comm_destroy(handle)
            if self.comm == "p2p":
                raise RuntimeError("comm='p2p' but the ranks cannot map each other's memory: %s" % err)
            self._comm = False
            return None
        self._comm = handle
        return handle

    def comm_name(self):
        return "p2p" if self._comm else "nccl"

    def check_status(self):
        """Raise if a peer GPU's data did not arrive in time during a distributed step (synchronises the stream)."""
        if not self._comm:
            return
        st = _lib.ctypes.c_int32(0)
        with torch.cuda.device(self.centers.device):
            _lib.call("acav_kmeans_comm_status", self._comm, _lib.ctypes.byref(st), _lib.stream_ptr(self.centers.device))
        if st.value != 0:
            raise RuntimeError("k-means distributed step: a peer GPU did not deliver its histogram / deltas / rows "
                               "within the spin limit; the centers of this rank are not valid")

    def _update(self, batch, counts_b, lr, dist, world, deltas=None):
        comm = self._peer_comm(dist, world) if world > 1 else None
        if comm is not None:                                                                        # :114-127, no NCCL
            b = batch.shape[0]
            with torch.cuda.device(self.centers.device):
                _lib.call("acav_kmeans_update_p2p", self._workspace(b), comm, _lib.ptr(batch, row_strided=True), b,
                          batch.stride(0), _lib.ptr(counts_b), lr, _lib.ptr(self.centers), _lib.ptr(self.counts),
                          _lib.ptr(self._fallback_dev), _lib.stream_ptr(self.centers.device))
        elif world > 1:
            dist.all_reduce(counts_b)                                                               # :114-115
            deltas = self._update_local(batch, counts_b, lr, deltas)                                # :116-123
            dist.all_reduce(deltas)                                                                 # :125-126
            self._apply_deltas(deltas)                                                              # :127
        else:
            self._update_fused(batch, counts_b, lr)                                                 # :116-127

    # -- CUDA-graph replay of the steady-state step ----------------------------------------------

    def _graphable(self, batch):
        return (self.graph in (True, "auto") and not self.in_warmup and batch.shape[0] > 0
                and not callable(self.lr) and type(self)._histogram is KMeans._histogram)

    def _graph_state(self, b, world):
        gs = self._gs
        if gs is None or gs["b"] != b or gs["world"] != world:
            dev = self._device()
            k, d = self.centers.shape
            self._workspace(b)
            gs = {
                "b": b, "world": world, "seen": set(), "graphs": collections.OrderedDict(), "i": 0,
                "best": torch.empty(b, dtype=torch.int64, device=dev),
                "mean": torch.empty(1, dtype=torch.float32, device=dev),
                "counts_b": torch.empty(k, dtype=torch.float32, device=dev),
                "deltas": torch.empty(k, d, dtype=torch.float32, device=dev) if world > 1 else None,
                "flags": torch.empty(k, dtype=torch.float32, device=dev),
                "thr": torch.empty(1, dtype=torch.float32, device=dev),
                "ring": torch.empty(_THR_RING, dtype=torch.float32).pin_memory(),
                "events": [None] * _THR_RING,
            }
            self._gs = gs
        return gs

    def _step_body(self, batch, gs, lr, distance, dist, world):
        """The step as a fixed sequence of C-ABI calls on the current stream: every by-value argument is constant for
        a given (batch, lr), the per-step threshold comes from device memory (gs['thr'])."""
        dev = self.centers.device
        k, d = self.centers.shape
        b = batch.shape[0]
        ws = self._workspace(b)
        with torch.cuda.device(dev):
            st = _lib.stream_ptr(dev)
            _lib.call("acav_kmeans_underused_flags", _lib.ptr(self.counts), k, _lib.ptr(gs["thr"]), _lib.ptr(gs["flags"]), st)
            _lib.call("acav_kmeans_assign", ws, _lib.ptr(batch, row_strided=True), b, batch.stride(0),
                      _lib.ptr(self.centers), _lib.ptr(gs["flags"]), 0.5, float(self.reinit[1]),
                      _lib.ptr(gs["best"]), None, _lib.ptr(gs["mean"]) if distance else None, None, self._mode(), st)
            _lib.call("acav_kmeans_histogram", ws, _lib.ptr(gs["best"]), b, _lib.ptr(gs["counts_b"]), st)
        self._update(batch, gs["counts_b"], lr, dist, world, deltas=gs["deltas"])

    def _add_graphed(self, batch, lr, distance, dist, world):
        dev = self._device()
        gs = self._graph_state(batch.shape[0], world)
        # this step's threshold: written to a pinned slot, copied to the device ahead of the replay
        slot = gs["i"] % _THR_RING
        gs["i"] += 1
        if gs["events"][slot] is not None:
            gs["events"][slot].synchronize()
        gs["ring"][slot] = self.underused_threshold()
        gs["thr"].copy_(gs["ring"][slot:slot + 1], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(dev))
        gs["events"][slot] = ev
        key = (batch.data_ptr(), batch.stride(0), lr, bool(distance), self._mode(), float(self.reinit[1]))
        g = gs["graphs"].get(key)
        if g is None and key not in gs["seen"]:
            # first time this batch address shows up: run eagerly (module loading, shared-memory attributes, workspace
            # allocation and the peer-memory setup must not happen inside a capture; batches that never come back --
            # fresh allocations of a streaming loader -- are never captured)
            if len(gs["seen"]) > 4 * _GRAPH_CACHE:
                gs["seen"].clear()
            gs["seen"].add(key)
            self._step_body(batch, gs, lr, distance, dist, world)
        else:
            if g is None:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._step_body(batch, gs, lr, distance, dist, world)
                gs["graphs"][key] = g
                if len(gs["graphs"]) > _GRAPH_CACHE:
                    gs["graphs"].popitem(last=False)
            else:
                gs["graphs"].move_to_end(key)
            g.replay()
        return gs["best"], (gs["mean"] if distance else None)
