"""``output.csv`` writer (reference subset_selection/code/save.py:6-44): headerless rows
``shard_name,filename,id,"[start, end]"`` APPENDED to the file; clips without metadata get
``id='-1', segment=[-1.0, -1.0]``."""
import csv
from collections import defaultdict
from pathlib import Path

from .. import hostio


def save_output(data, metas, out_path, name='', sharded_meta=True):
    out_path = Path(out_path)
    out_path.parent.mkdir(exist_ok=True, parents=True)
    rows, keys = {}, []
    for row in data:
        fname = hostio.file_stem(row['filename'])
        meta = None
        if sharded_meta:
            meta = metas.get(row['shard_name'], {}).get(fname)
        else:
            meta = metas.get(fname)
        if meta is None:
            meta = {'id': '-1', 'segment': [-1.0, -1.0]}
        rows[fname] = {**row, **meta}
        keys.append(fname)
    headers = ['shard_name', 'filename', 'id', 'segment']
    out_path = out_path.parent / (name + out_path.name)
    count = 0
    with open(out_path, 'a+') as f:
        writer = csv.writer(f)
        for key in keys:
            writer.writerow([rows[key][h] for h in headers])
            count += 1
    return out_path, count


def merge_csvs(ins, out):
    """save.py:85-93 -- append the given CSVs (sorted by path) to `out`; returns the number of lines added."""
    count = 0
    with open(out, 'a+') as out_f:
        for in_file in sorted(ins):
            with open(in_file, 'r') as in_f:
                for line in in_f:
                    out_f.write(line)
                    count += 1
    return count


def group_cache_paths(paths):
    """save.py:96-103 -- caches of one run share the prefix ``cache_<parent pid>``."""
    groups = defaultdict(list)
    for path in paths:
        groups['_'.join(Path(path).stem.split('_')[:2])].append(Path(path))
    return {k: sorted(v) for k, v in groups.items()}


def merge_all_csvs(args):
    """save.py:106-121 -- ``reduce_csvs``: append every cached per-chunk CSV to the output CSV."""
    out_path = Path(args.data.output.path)
    cache_dir = out_path.parent / 'caches'
    groups = group_cache_paths(list(cache_dir.glob('cache_*_*_{}'.format(out_path.name))))
    for key in sorted(groups.keys()):
        print('processing cache set {}'.format(key))
        print("merging csvs")
        counts = merge_csvs(groups[key], out_path)
        if args.verbose:
            print("Saved Results: added {} lines to {}".format(counts, out_path))
