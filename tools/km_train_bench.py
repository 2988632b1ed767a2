"""k-means in the state the operator actually reaches: train from the reference init (torch.rand*1e-5, random-assignment
warm-up, lr schedule of run_clustering.py:168) for whole epochs over a resident shard, print per epoch what the model looks
like (clusters in use, lr fallbacks, rows needing the exact re-check), then time steady-state steps eagerly and replayed from
CUDA graphs.

    python tools/km_train_bench.py [--rows 1250000 --b 8192 --d 2048 --k 1024 --epochs 2 --steps 100] [--profile]
    ncu --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file out.csv python tools/km_train_bench.py --profile
"""
import argparse
import json
import sys
import time
import types

import torch

sys.path.insert(0, ".")
from acav100m_b200 import _lib, synth
from acav100m_b200.clustering import KMeans

p = argparse.ArgumentParser()
p.add_argument("--rows", type=int, default=1_250_000)
p.add_argument("--b", type=int, default=8192)
p.add_argument("--d", type=int, default=2048)
p.add_argument("--k", type=int, default=1024)
p.add_argument("--epochs", type=int, default=2)
p.add_argument("--steps", type=int, default=100)
p.add_argument("--mode", default="auto")
p.add_argument("--profile", action="store_true", help="mark 3 graph-replayed steps for ncu and exit")
a = p.parse_args()

dev = torch.device("cuda", 0)
x = synth.gaussian_mixture_torch(a.rows, a.d, a.k, 2003, dev, means_seed=1003)          # bench.py's rank-0 shard
nb = a.rows // a.b
kargs = types.SimpleNamespace(computation=types.SimpleNamespace(device="cuda", num_gpus=1))


def snapshot(km, tag):
    xb = x[:a.b]
    ws = km._workspace(a.b)
    best = torch.empty(a.b, dtype=torch.int64, device=dev)
    nref = torch.zeros(2, dtype=torch.int32, device=dev)
    if not km.in_warmup:
        _lib.call("acav_kmeans_assign", ws, _lib.ptr(xb), a.b, a.d, _lib.ptr(km.centers), _lib.ptr(km.counts),
                  km.underused_threshold(), float(km.reinit[1]), _lib.ptr(best), None, None, _lib.ptr(nref), km._mode(),
                  _lib.stream_ptr(dev))
    hist = torch.bincount(best, minlength=a.k)
    return {"tag": tag, "count": km.count, "lr_fallbacks": km.fallback, "clusters_in_use_batch": int((hist > 0).sum()),
            "max_rows_per_centroid": int(hist.max()), "rows_rechecked_on_candidates": int(nref[0]),
            "rows_through_full_exact_kernel": int(nref[1]), "underused": int((km.counts < km.underused_threshold()).sum())}


def time_steps(km, steps, off):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0.record()
    for i in range(steps):
        j = (off + i) % nb
        km.add(x[j * a.b:(j + 1) * a.b], sync=False, distance=False)
    e1.record()
    t_issue = time.perf_counter() - t0
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, t_issue * 1e3 / steps


out = {"shape": vars(a)}
for graph in ("auto", False):
    torch.manual_seed(1003)
    km = KMeans(kargs, a.d, a.k, assign_mode=a.mode, warmup_rng="cuda", graph=graph)
    km.to(dev)
    log = []
    t0 = time.perf_counter()
    for epoch in range(a.epochs):
        km.lr = 0.1 ** (2 + epoch // 5)
        for j in range(nb):
            km.add(x[j * a.b:(j + 1) * a.b], sync=False, distance=False)
            if epoch == 0 and j in (3, 10, 30):
                log.append(snapshot(km, "epoch 0 step %d" % j))
        torch.cuda.synchronize()
        log.append(snapshot(km, "after epoch %d (%.2f s)" % (epoch, time.perf_counter() - t0)))
    if a.profile:
        time_steps(km, 3, 0)
        torch.cuda.profiler.start()
        time_steps(km, 3, 3)
        torch.cuda.profiler.stop()
        print(json.dumps(log))
        sys.exit(0)
    time_steps(km, 5, 0)
    ms, issue = time_steps(km, a.steps, 5)
    key = "graph" if graph else "eager"
    out[key] = {"ms_per_step": ms, "host_issue_ms_per_step": issue, "iter_per_sec": 1e3 / ms,
                "graphs_captured": len(km._gs["graphs"]) if km._gs else 0, "launches_per_step": km.launches_per_step()}
    if graph:
        out["training_log"] = log
        # skewed state for comparison: the first steps after warm-up of a fresh model
        torch.manual_seed(1003)
        km2 = KMeans(kargs, a.d, a.k, assign_mode=a.mode, warmup_rng="cuda", graph=False)
        km2.to(dev)
        km2.lr = 1e-2
        settle = -(-10 * a.k // a.b) + 8
        for j in range(settle):
            km2.add(x[j * a.b:(j + 1) * a.b], sync=False, distance=False)
        ms2, _ = time_steps(km2, 16, settle)
        out["early_state_eager_ms_per_step"] = ms2
        out["early_state"] = snapshot(km2, "early")
    centers_key = "centers_" + key
    out[centers_key] = float(km.centers.double().abs().sum())
print(json.dumps(out))
