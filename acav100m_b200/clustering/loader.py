"""Parallel feature-shard loader: the reference's DataLoader worker processes (clustering/code/data/clustering.py:17-33,
116-197) rebuilt so that the workers do the expensive part -- unpickling a shard's Python rows and collating them -- and
hand the parent whole arrays through SHARED MEMORY instead of pickling 24 KB per clip back through a pipe.

    parent:  ShardLoader(paths, batch_size, drop_last, workers=N)  ->  iterates the SAME collated batches, in the same
             order, as data.batches(paths, batch_size, drop_last) (single-worker order: shard order, row order)
    worker:  one shard at a time: pickle.load -> one float32 [n, d] array per (extractor, layer), written straight into
             one multiprocessing.shared_memory block; the row metadata (filenames, shard names, sizes) travels through
             the result pipe (a few bytes per clip)

The parent maps the block (zero copy), optionally page-locks it (cudaHostRegister) so that the H2D copies of the batches
cut from it are asynchronous DMA, slices batches out of it (a batch that straddles two shards is the only copy) and
unlinks the block when the shard is used up.  Workers never touch CUDA.
"""
import collections
import multiprocessing as mp
import pickle
from multiprocessing import shared_memory
from pathlib import Path

import numpy as np

FEATURE_KEYS = ('video_features', 'audio_features')


def _layer_keys(array):
    if isinstance(array, dict):
        return list(array.keys())
    if isinstance(array, (list, tuple)):
        return ['layer_{}'.format(i) for i in range(len(array))]
    return None


def _get_layer(array, layer):
    if isinstance(array, dict):
        return array[layer]
    return array[int(layer.split('_')[-1])]


def _plan(rows):
    """[(feature key, index, model key 'EXTRACTOR/dataset', layer or None, row shape)] of a shard, from its first row."""
    pivot = rows[0]
    plan = []
    for key in pivot.keys():
        if key in FEATURE_KEYS:
            for i, feat in enumerate(pivot[key]):
                name = '/'.join((feat['extractor_name'], feat['dataset']))
                layers = _layer_keys(feat['array'])
                if layers is None:
                    plan.append((key, i, name, None, tuple(np.asarray(feat['array']).shape)))
                else:
                    for layer in layers:
                        plan.append((key, i, name, layer, tuple(np.asarray(_get_layer(feat['array'], layer)).shape)))
    return plan


def _load_shard(path):
    """Worker: -> (shm name, nbytes, n rows, [(model key, layer, shape, byte offset)], {meta key: list}) or None."""
    try:
        with open(path, 'rb') as f:
            rows = pickle.load(f)
    except Exception as e:  # noqa: BLE001 -- same tolerance as the reference (data/clustering.py:167-182)
        print('Exception in shard loading: {} ({})'.format(Path(path).stem, e))
        return None
    n = len(rows)
    if n == 0:
        return None
    plan = _plan(rows)
    offsets, total = [], 0
    for _, _, _, _, shape in plan:
        offsets.append(total)
        total += int(np.prod(shape, dtype=np.int64)) * n * 4
        total = (total + 255) // 256 * 256
    shm = shared_memory.SharedMemory(create=True, size=max(total, 256))
    try:
        from multiprocessing import resource_tracker
        resource_tracker.unregister(shm._name, 'shared_memory')        # the parent owns the block's lifetime
    except Exception:  # noqa: BLE001
        pass
    layout = []
    for (key, i, name, layer, shape), off in zip(plan, offsets):
        out = np.ndarray((n,) + shape, dtype=np.float32, buffer=shm.buf, offset=off)
        # np.stack(..., out=) writes the rows straight into the block (3x faster than assigning row by row)
        if layer is None:
            np.stack([np.asarray(row[key][i]['array'], dtype=np.float32) for row in rows], out=out)
        else:
            first = rows[0][key][i]['array']
            lk = layer if isinstance(first, dict) else int(layer.split('_')[-1])
            kind = type(first)
            np.stack([np.asarray(row[key][i]['array'][lk] if type(row[key][i]['array']) is kind
                                 else _get_layer(row[key][i]['array'], layer), dtype=np.float32) for row in rows], out=out)
        layout.append((name, layer, shape, off))
    meta = {k: [row[k] for row in rows] for k in rows[0].keys() if k not in FEATURE_KEYS}
    name = shm.name
    shm.close()
    return name, total, n, layout, meta


class _Shard:
    """A loaded shard in the parent: arrays are views of the shared-memory block."""

    def __init__(self, result, pin):
        name, nbytes, n, layout, meta = result
        self.shm = shared_memory.SharedMemory(name=name)
        self.n, self.meta, self.pos = n, meta, 0
        self.pinned_ptr = None
        self.arrays = collections.OrderedDict()
        for model, layer, shape, off in layout:
            # np.frombuffer keeps a buffer EXPORT on the mapping for as long as any view of the array lives (np.ndarray(
            # buffer=...) does not: closing the block under such an array leaves it dangling)
            count = n * int(np.prod(shape, dtype=np.int64))
            arr = np.frombuffer(self.shm.buf, dtype=np.float32, count=count, offset=off).reshape((n,) + tuple(shape))
            if layer is None:
                self.arrays[model] = arr
            else:
                self.arrays.setdefault(model, collections.OrderedDict())[layer] = arr
        if pin and nbytes > 0:
            self._pin(nbytes)

    def _pin(self, nbytes):
        try:
            import torch
            if not torch.cuda.is_available():
                return
            ptr = np.frombuffer(self.shm.buf, dtype=np.uint8, count=1).ctypes.data
            if int(torch.cuda.cudart().cudaHostRegister(ptr, nbytes, 0)) == 0:
                self.pinned_ptr = ptr
        except Exception:  # noqa: BLE001 -- pinning is an optimisation
            self.pinned_ptr = None

    def release(self):
        if self.pinned_ptr is not None:
            try:
                import torch
                torch.cuda.synchronize()                       # batches cut from the block may still be in flight
                torch.cuda.cudart().cudaHostUnregister(self.pinned_ptr)
            except Exception:  # noqa: BLE001
                pass
            self.pinned_ptr = None
        self.arrays = None
        try:
            self.shm.unlink()                                  # the name goes now; the pages when the last view does
        except Exception:  # noqa: BLE001
            pass
        try:
            self.shm.close()
        except BufferError:                                    # a batch cut from this block is still alive somewhere
            pass


def _cut(arrays, lo, hi):
    return {m: ({layer: a[lo:hi] for layer, a in v.items()} if isinstance(v, dict) else v[lo:hi]) for m, v in arrays.items()}


def _concat(parts):
    first = parts[0]
    out = {}
    for m, v in first.items():
        if isinstance(v, dict):
            out[m] = {layer: np.concatenate([p[m][layer] for p in parts]) for layer in v}
        else:
            out[m] = np.concatenate([p[m] for p in parts])
    return out


class ShardLoader:
    """Iterable over collated batches (numpy arrays, `torch.from_numpy`-ready) -- see the module docstring.
    `workers` <= 0 loads in-process (no shared memory).  `hold` shards stay mapped after they are used up (their batches
    may still be feeding asynchronous H2D copies); older ones are released."""

    def __init__(self, shard_paths, batch_size, drop_last, workers=4, prefetch=None, pin=False, as_torch=True, hold=2):
        self.paths = [str(p) for p in shard_paths]
        self.batch_size, self.drop_last = int(batch_size), drop_last
        self.workers = int(workers)
        self.prefetch = prefetch if prefetch is not None else max(2, 2 * self.workers)
        self.pin, self.as_torch, self.hold = pin, as_torch, hold

    def _results(self):
        if self.workers <= 0:
            for p in self.paths:
                yield _load_shard(p)
            return
        ctx = mp.get_context('fork')
        with ctx.Pool(self.workers) as pool:
            pending = collections.deque()
            it = iter(self.paths)
            for p in it:
                pending.append(pool.apply_async(_load_shard, (p,)))
                if len(pending) >= self.prefetch:
                    break
            while pending:
                res = pending.popleft().get()
                nxt = next(it, None)
                if nxt is not None:
                    pending.append(pool.apply_async(_load_shard, (nxt,)))
                yield res

    def _finish(self, feats, meta):
        from .. import hostio
        batch = {}
        if self.as_torch:
            import torch
            for m, v in feats.items():
                batch[m] = ({layer: torch.from_numpy(a) for layer, a in v.items()} if isinstance(v, dict)
                            else torch.from_numpy(v))
        else:
            batch.update(feats)
        batch.update(meta)
        batch['idx'] = [hostio.file_stem(f) for f in meta['filename']]
        return batch

    def __iter__(self):
        b = self.batch_size
        parts, metas, have = [], [], 0          # pieces of the batch being assembled
        used = collections.deque()
        try:
            for res in self._results():
                if res is None:
                    continue
                shard = _Shard(res, self.pin)
                while shard.pos < shard.n:
                    take = min(b - have, shard.n - shard.pos)
                    parts.append(_cut(shard.arrays, shard.pos, shard.pos + take))
                    metas.append({k: v[shard.pos:shard.pos + take] for k, v in shard.meta.items()})
                    shard.pos += take
                    have += take
                    if have == b:
                        feats = parts[0] if len(parts) == 1 else _concat(parts)
                        meta = {k: sum((m[k] for m in metas), []) for k in metas[0]}
                        yield self._finish(feats, meta)
                        parts, metas, have = [], [], 0
                if parts:                       # the tail of this shard waits for the next one: copy it out of the block
                    parts = [_concat(parts)] if len(parts) > 1 else [{m: ({l: a.copy() for l, a in v.items()}
                                                                          if isinstance(v, dict) else v.copy())
                                                                      for m, v in parts[0].items()}]
                    metas = [{k: sum((m[k] for m in metas), []) for k in metas[0]}]
                used.append(shard)
                while len(used) > self.hold:
                    used.popleft().release()
            if have and not self.drop_last:
                feats = parts[0] if len(parts) == 1 else _concat(parts)
                meta = {k: sum((m[k] for m in metas), []) for k in metas[0]}
                yield self._finish(feats, meta)
        finally:
            while used:
                used.popleft().release()
