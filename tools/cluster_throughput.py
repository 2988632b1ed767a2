"""Clips per second through `cli cluster` (train + assign) on synthetic feature shards in the reference's format, and of
the shard loader alone for 0 .. N worker processes:
    python tools/cluster_throughput.py [--shards 16 --clips 2000 --batch 1024 --k 256] > gpurun_out/cluster_throughput.json"""
import argparse
import json
import os
import shutil
import sys
import tempfile
import time
from pathlib import Path

import torch

sys.path.insert(0, ".")
from acav100m_b200.clustering import cli as ccli, data as cdata, loader      # noqa: E402
from tests.shard_fixtures import write_feature_shards                         # noqa: E402

p = argparse.ArgumentParser()
p.add_argument("--shards", type=int, default=16)
p.add_argument("--clips", type=int, default=2000)
p.add_argument("--batch", type=int, default=1024)
p.add_argument("--k", type=int, default=256)
a = p.parse_args()

root = Path(tempfile.mkdtemp(prefix="acav_thr_"))
out = {"shards": a.shards, "clips_per_shard": a.clips, "batch": a.batch, "k": a.k, "cores": os.cpu_count(),
       "bytes_per_clip": 5944 * 4}
try:
    feat_dir, meta_dir = write_feature_shards(root / "data", n_shards=a.shards, clips_per_shard=a.clips, seed=1)
    glob = str(feat_dir / ("shard-{000000..%06d}.pkl" % (a.shards - 1)))
    paths = cdata.expand_shards(glob)
    total = a.shards * a.clips
    rates = {}
    t0 = time.perf_counter()
    n = sum(len(b["idx"]) for b in cdata.batches(paths, a.batch, True))
    rates["in_process_reader"] = n / (time.perf_counter() - t0)
    for w in (1, 2, 4, 8, 16, 32):
        if w >= (os.cpu_count() or 2):
            break
        t0 = time.perf_counter()
        n = sum(len(b["idx"]) for b in loader.ShardLoader(paths, a.batch, True, workers=w))
        rates["loader_workers_%d" % w] = n / (time.perf_counter() - t0)
    out["loader_clips_per_sec"] = rates
    for tag, epochs in (("one_epoch", 1), ("three_epochs", 3)):
        clusters = root / ("clusters_" + tag)
        torch.manual_seed(1)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ccli.main(["cluster", "--feature_path=" + glob, "--out_path=" + str(clusters), "--meta_path=" + str(meta_dir),
                   "--clustering.ncentroids=%d" % a.k, "--data.batch_size=%d" % a.batch, "--clustering.epochs=%d" % epochs,
                   "--computation.num_gpus=1"])
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        out["cli_cluster_" + tag] = {"seconds": dt, "clips_trained_plus_assigned": total * (epochs + 1),
                                     "clips_per_sec": total * (epochs + 1) / dt}
finally:
    shutil.rmtree(root, ignore_errors=True)
print(json.dumps(out))
