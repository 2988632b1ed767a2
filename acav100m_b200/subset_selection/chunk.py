"""Chunked selection over many shards, one process per GPU (reference subset_selection/code/chunk.py:21-226).

The shard list is cut into chunks of ``--chunk_size`` shards; chunks are dealt to ``--computation.num_gpus``
processes in contiguous runs (``split_chunks``); every chunk is an independent greedy selection with quota
``ceil(subset.size / num_chunks)`` (or ``subset.ratio`` of the chunk); each chunk's result is cached under
``<out dir>/caches/`` either as its own CSV (``--save_cache_as_csvs``, default) or as a pickle
``{'res', 'metas'}``; ``reduce`` / ``reduce_csvs`` / ``reduce_pkls`` append the caches to the output CSV.
This is the reference's own multi-GPU mode: replicas of the single-GPU operator, no data-path collective.
"""
import concurrent.futures
import copy
import math
import os
from functools import reduce
from pathlib import Path

from .. import hostio
from .dataloader import load_data
from .run import run_partition
from .save import group_cache_paths, merge_csvs, save_output


def get_chunks(li, n):
    """utils.py:76-79 -- split into chunks of size n."""
    for i in range(0, len(li), n):
        yield li[i: i + n]


def split_chunks(li, m):
    """utils.py:82-85 -- split into m contiguous runs."""
    n = math.ceil(len(li) / m)
    return get_chunks(li, n)


def plan_chunks(args):
    """chunk.py:23-48 -> (chunk_args, per-process lists of (chunk number, shard paths), num_chunks)."""
    shard_paths = sorted(hostio.braceexpand(str(args.data.path)))
    shard_paths = [p for p in shard_paths if Path(p).is_file()]
    chunks = list(get_chunks(shard_paths, args.chunk_size))
    num_chunks = len(chunks)
    if num_chunks == 0:
        raise FileNotFoundError("no cluster shards match {}".format(args.data.path))
    if args.computation.num_gpus > num_chunks:
        print("num_gpus ({}) exceeds num_chunks ({})".format(args.computation.num_gpus, num_chunks))
        print("thresholding num_gpus to be equal to num_chunks")
        args.computation.num_gpus = num_chunks
    nodes_chunks = list(split_chunks(list(enumerate(chunks)), args.computation.num_gpus))
    chunk_args = copy.deepcopy(args)
    if isinstance(chunk_args.subset.size, int):
        chunk_args.subset.size = math.ceil(chunk_args.subset.size / num_chunks)
    chunk_args.computation.num_workers = round(chunk_args.computation.num_workers / chunk_args.computation.num_gpus)
    return chunk_args, nodes_chunks, num_chunks


def run_chunks(args):
    """chunk.py:21-53."""
    args.parent_pid = str(os.getpid())
    chunk_args, nodes_chunks, num_chunks = plan_chunks(args)
    print("running {} chunks in {} gpus".format(num_chunks, chunk_args.computation.num_gpus))
    cfg = (chunk_args, nodes_chunks, num_chunks)
    if chunk_args.computation.num_gpus <= 1:
        run_chunks_node(0, cfg)
        return
    import torch.multiprocessing as mp
    mp.spawn(run_chunks_node, args=(cfg,), nprocs=chunk_args.computation.num_gpus, join=True)


def run_chunks_node(node_rank, cfg):
    """chunk.py:113-132 -- one process: its run of chunks on GPU `node_rank`."""
    args, nodes_chunks, num_chunks = cfg
    args = copy.deepcopy(args)
    args.node_rank = node_rank
    if node_rank >= len(nodes_chunks):
        return
    if args.computation.device == 'cuda':
        import torch
        torch.cuda.set_device(node_rank % max(torch.cuda.device_count(), 1))
    chunks = nodes_chunks[node_rank]
    loaded = None
    with concurrent.futures.ThreadPoolExecutor(max_workers=1) as pool:
        for i, (num, chunk) in enumerate(chunks):
            # chunk.py:190-226 (load_async): the next chunk's shards are unpickled while this one is selected
            data = loaded.result() if loaded is not None else _load_chunk(args, chunk)
            loaded = None
            if args.computation.load_async and i + 1 < len(chunks):
                loaded = pool.submit(_load_chunk, args, chunks[i + 1][1])
            print("running chunk {}".format(num))
            res, metas = _select_chunk(args, data)
            name = "cache_{}_{}_{}".format(args.parent_pid, node_rank, i)
            if args.save_cache_as_csvs:
                _reduce_single_cache(args, name, res, metas)
            else:
                save_chunk_cache(args, node_rank, i, res, metas)


def _load_chunk(args, chunk):
    return load_data(list(chunk), args.data.meta.path, verbose=False)


def _select_chunk(args, data):
    """run.py:13-48 on one chunk; chunk.py:151-153 keeps the first partition's result."""
    partitions, metas = data
    results = [run_partition(args, partitions[k], metas) for k in sorted(partitions.keys())]
    return results[0], metas


def save_chunk_cache(args, rank, i, res, metas):
    """chunk.py:135-146."""
    cache_dir = Path(args.data.output.path).parent / 'caches'
    cache_dir.mkdir(parents=True, exist_ok=True)
    name = "cache_{}_{}_{}.pkl".format(args.parent_pid, rank, i)
    hostio.dump_pickle({'res': res, 'metas': metas}, str(cache_dir / name))


def _reduce_single_cache(args, k, res, metas):
    """chunk.py:91-110.  `metas` is {shard stem: {clip stem: row}}; the reference flattens it to
    {clip stem: row} here and looks rows up without the shard level."""
    flat = reduce(lambda x, y: {**x, **y}, metas.values()) if len(metas) else {}
    if isinstance(args.subset.size, int):
        res = res[:args.subset.size]         # the reference reads `args.subset_size` here (an AttributeError on its
        #                                      Munch tree); the evident intent is the per-chunk quota
    print("saving cache ({}), subset size: {}".format(k, len(res)))
    out = Path(args.data.output.path)
    possible_out_path = out.parent / 'caches' / out.name
    return save_output(res, flat, possible_out_path, k + '_', sharded_meta=False)


def reduce_single_cache(args, path):
    """chunk.py:82-88."""
    print("loading cache ({})".format(Path(path).stem))
    cache = hostio.load_pickle(path)
    return _reduce_single_cache(args, Path(path).stem, cache['res'], cache['metas'])


def reduce_all_pkls(args):
    """chunk.py:56-79 -- pickled chunk caches -> per-cache CSVs -> appended to the output CSV."""
    cache_dir = Path(args.data.output.path).parent / 'caches'
    groups = group_cache_paths(list(cache_dir.glob('cache_*_*.pkl')))
    for key in sorted(groups.keys()):
        print('processing cache set {}'.format(key))
        out_paths = [reduce_single_cache(args, p)[0] for p in groups[key]]
        if len(out_paths) == 0:
            print("No files saved")
        print("merging csvs")
        counts = merge_csvs(sorted(out_paths), args.data.output.path)
        if args.verbose:
            print("Saved Results: added {} lines to {}".format(counts, args.data.output.path))
