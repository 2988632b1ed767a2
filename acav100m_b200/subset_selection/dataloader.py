"""Cluster-shard loading for the selection stage (reference subset_selection/code/dataloader.py:17-255).

Reads the pickles written by the clustering stage (``[epoch_e_]shard-NNNNNN.pkl``, rows with
``audio_assignments`` / ``video_assignments``), groups shards into partitions by the ``log_*.json`` files
next to them (newest log wins, shards without a log form partition -1), joins the per-shard metadata
JSON (``{filename, id, segment}``) and flattens each partition to ``assignments int64 [V, D]`` whose
columns are the SORTED ``(model_key, layer)`` tuples (dataloader.py:44-53).
"""
from collections import defaultdict
from pathlib import Path

import numpy as np

from .. import hostio


def format_row(row):
    """dataloader.py:17-36 (the dict-of-layers branch is the one that works in the reference; plain
    arrays are filed under layer 'model')."""
    res = {}
    for feature_name in ('audio_assignments', 'video_assignments'):
        for feature in row[feature_name]:
            array = feature['array']
            if isinstance(array, dict):
                for layer, value in array.items():
                    res[(feature['model_key'], layer)] = value
            elif isinstance(array, (list, tuple)):
                for i, value in enumerate(array):
                    res[(feature['model_key'], 'layer_{}'.format(i))] = value
            else:
                res[(feature['model_key'], 'model')] = array
    return row['filename'], row['shard_name'], res


def preprocess(data, columns=None):
    """dataloader.py:56-69 -> (assignments [V, D] int64, shard_names, filenames, clustering_types)."""
    filenames, shard_names, rows = zip(*(format_row(r) for r in data))
    clustering_types = sorted(rows[0].keys())
    if columns is not None:
        columns = [tuple(c) for c in columns]
        missing = [c for c in columns if c not in clustering_types]
        assert not missing, "clustering.columns not present in the shards: {}".format(missing)
        clustering_types = columns
    assignments = np.array([[int(r[k]) for k in clustering_types] for r in rows], dtype=np.int64)
    return assignments, shard_names, filenames, clustering_types


def load_partitions(shards_dir):
    """dataloader.py:72-83 -- shard name -> partition id, newer logs override older ones."""
    log_paths = sorted(Path(shards_dir).glob('log_*.json'), key=lambda x: str(x).split('.')[-2].split('_')[-1])
    partitions = {}
    for i, log_path in enumerate(log_paths):
        for shard in hostio.load_json(log_path)['shards']:
            partitions[shard] = i
    return partitions


def load_metas(shard_paths, metas_path):
    """dataloader.py:206-255 -- {shard stem: {clip stem: meta row}}."""
    metas = {}
    for shard_path in shard_paths:
        meta_path = Path(metas_path) / "{}.json".format(shard_path.stem)
        if meta_path.is_file():
            metas[shard_path.stem] = {Path(r['filename']).stem: r for r in hostio.load_json(meta_path)}
    return metas


def load_data(shard_paths, metas_path, verbose=False):
    """dataloader.py:152-203 -> ({partition id: [rows]}, metas)."""
    if not isinstance(shard_paths, list):
        shard_paths = hostio.braceexpand(str(shard_paths))
    partitions = load_partitions(Path(shard_paths[0]).parent)
    shard_paths = sorted(p for p in (Path(s) for s in shard_paths) if p.is_file())
    partitioned = defaultdict(list)
    for shard_path in shard_paths:
        partitioned[partitions.get(shard_path.stem, -1)] += hostio.load_pickle(shard_path)
    if verbose:
        print("num_shards: {} (dataset_size per partition: {})".format(
            len(shard_paths), {k: len(v) for k, v in partitioned.items()}))
    return partitioned, load_metas(shard_paths, metas_path)
