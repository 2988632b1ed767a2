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


def _column_plan(row, clustering_types):
    """Where each clustering of `clustering_types` sits inside a row: (feature list name, index in that list, model key,
    key inside 'array') -- lets `preprocess` read the ten ids of a row without building a dict per row."""
    plan = []
    for model_key, layer in clustering_types:
        hit = None
        for feature_name in ('audio_assignments', 'video_assignments'):
            for pos, feature in enumerate(row[feature_name]):
                if feature['model_key'] != model_key:
                    continue
                array = feature['array']
                if isinstance(array, dict):
                    if layer in array:
                        hit = (feature_name, pos, model_key, layer)
                elif isinstance(array, (list, tuple)):
                    if layer.startswith('layer_') and layer[6:].isdigit() and int(layer[6:]) < len(array):
                        hit = (feature_name, pos, model_key, int(layer[6:]))
                elif layer == 'model':
                    hit = (feature_name, pos, model_key, None)
        if hit is None:
            return None
        plan.append(hit)
    return plan


def _row_ids(row, plan):
    """The row's cluster ids in plan order, or None when the row is laid out differently from the first one."""
    out = []
    try:
        for feature_name, pos, model_key, key in plan:
            feature = row[feature_name][pos]
            if feature['model_key'] != model_key:
                return None
            out.append(feature['array'] if key is None else feature['array'][key])
    except (IndexError, KeyError, TypeError):
        return None
    return out


def preprocess(data, columns=None):
    """dataloader.py:56-69 -> (assignments [V, D] int64, shard_names, filenames, clustering_types).  Same result as
    formatting every row into a {(model_key, layer): id} dict (`format_row`); rows laid out like the first one -- all of
    them, in shards written by the clustering stage -- are read through a precomputed column plan instead."""
    clustering_types = sorted(format_row(data[0])[2].keys())
    if columns is not None:
        columns = [tuple(c) for c in columns]
        missing = [c for c in columns if c not in clustering_types]
        assert not missing, "clustering.columns not present in the shards: {}".format(missing)
        clustering_types = columns
    plan = _column_plan(data[0], clustering_types)
    flat = []
    for row in data:
        ids = _row_ids(row, plan) if plan is not None else None
        if ids is None:
            formatted = format_row(row)[2]
            ids = [formatted[k] for k in clustering_types]
        flat.extend(ids)
    assignments = np.array(flat, dtype=np.int64).reshape(len(data), len(clustering_types))
    filenames = tuple(r['filename'] for r in data)
    shard_names = tuple(r['shard_name'] for r in data)
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
            metas[shard_path.stem] = {hostio.file_stem(r['filename']): r for r in hostio.load_json(meta_path)}
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
