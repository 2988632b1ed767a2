"""Feature-shard reader and collate for the clustering stage.

On-disk format (written by feature_extraction): ``shard-NNNNNN.pkl`` = list of
rows ``{'filename', 'shard_name', 'shard_size', 'video_features': [feat..], 'audio_features': [feat..]}``
with ``feat = {'model_key', 'extractor_name', 'dataset', 'array': {'layer_i': float32[d_i]} | float32[d]}``.
Collated batches have the layout of the reference's ``collate_features``
(clustering/code/data/clustering.py:78-113): ``{'EXTRACTOR/dataset': {layer: FloatTensor[b, d]} | Tensor,
'filename': [...], 'shard_name': [...], 'shard_size': [...], 'idx': [stem, ...]}``.

The reference streams rows through DataLoader worker processes and webdataset's ResizedDataset; here
rows are read in-process in shard order (the batch sequence of a single-worker run).  Shard -> rank
selection for the assignment pass is the reference's ``shards[rank::world]``
(mps/distributed.py:438-439); training splits every global batch contiguously over the ranks.
"""
from pathlib import Path

import numpy as np
import torch

from .. import hostio

FEATURE_KEYS = ('video_features', 'audio_features')


def expand_shards(path):
    paths = [Path(p) for p in hostio.braceexpand(str(path))]
    return [p for p in paths if p.is_file()]


def iter_rows(shard_paths):
    """Rows of all shards in order; unreadable shards are reported and skipped
    (data/clustering.py:167-182)."""
    for path in shard_paths:
        try:
            rows = hostio.load_pickle(path)
        except Exception as e:  # noqa: BLE001 -- same tolerance as the reference
            print('Exception in shard loading: {} ({})'.format(Path(path).stem, e))
            continue
        for row in rows:
            yield row


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


def _stack_layer(arrays, layer):
    """float32 [b, ...] of one layer over the rows' 'array' containers: rows written straight into a preallocated
    buffer (np.stack builds a view object per row first: 1.6x slower on 10^4 clips)."""
    first = np.asarray(_get_layer(arrays[0], layer), dtype=np.float32)
    out = np.empty((len(arrays),) + first.shape, dtype=np.float32)
    kind = type(arrays[0])
    key = layer if isinstance(arrays[0], dict) else int(layer.split('_')[-1])
    for j, arr in enumerate(arrays):
        out[j] = arr[key] if type(arr) is kind else _get_layer(arr, layer)
    return out


def collate_features(rows):
    pivot = rows[0]
    res = {}
    for key in pivot.keys():
        if key in FEATURE_KEYS:
            for i, feat in enumerate(pivot[key]):
                layers = _layer_keys(feat['array'])
                if layers is not None:
                    arrays = [r[key][i]['array'] for r in rows]
                    feature = {layer: torch.from_numpy(_stack_layer(arrays, layer)) for layer in layers}
                else:
                    feature = torch.from_numpy(np.stack(
                        [np.asarray(r[key][i]['array'], dtype=np.float32) for r in rows]))
                res['/'.join((feat['extractor_name'], feat['dataset']))] = feature
        else:
            res[key] = [r[key] for r in rows]
    res['idx'] = [hostio.file_stem(r['filename']) for r in rows]
    return res


def batches(shard_paths, batch_size, drop_last):
    buf = []
    for row in iter_rows(shard_paths):
        buf.append(row)
        if len(buf) == batch_size:
            yield collate_features(buf)
            buf = []
    if buf and not drop_last:
        yield collate_features(buf)


def rank_slice(batch, rank, world):
    """Contiguous 1/world slice of a collated global batch (per-rank batch = batch_size / world,
    data/clustering.py:25)."""
    if world == 1:
        return batch
    n = len(batch['idx'])
    per = n // world
    lo, hi = rank * per, (rank + 1) * per

    def cut(v):
        if isinstance(v, dict):
            return {k: cut(x) for k, x in v.items()}
        return v[lo:hi]

    return {k: cut(v) for k, v in batch.items()}
