"""``output.csv`` writer (reference subset_selection/code/save.py:6-44): headerless rows
``shard_name,filename,id,"[start, end]"`` APPENDED to the file; clips without metadata get
``id='-1', segment=[-1.0, -1.0]``."""
import csv
from pathlib import Path


def save_output(data, metas, out_path, name='', sharded_meta=True):
    out_path = Path(out_path)
    out_path.parent.mkdir(exist_ok=True, parents=True)
    rows, keys = {}, []
    for row in data:
        fname = Path(row['filename']).stem
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
