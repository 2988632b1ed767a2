"""Steady-state k-means step for profiling: train unprofiled, then mark a few steps for ncu.

    ncu --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file out.csv \
        python tools/km_steady.py [--b 8192 --d 2048 --k 1024 --settle 60 --steps 3]
"""
import argparse
import sys
import types

import torch

sys.path.insert(0, ".")
from acav100m_b200 import synth
from acav100m_b200.clustering import KMeans

p = argparse.ArgumentParser()
p.add_argument("--b", type=int, default=8192)
p.add_argument("--d", type=int, default=2048)
p.add_argument("--k", type=int, default=1024)
p.add_argument("--rows", type=int, default=400_000)
p.add_argument("--settle", type=int, default=60)
p.add_argument("--steps", type=int, default=3)
p.add_argument("--mode", default="auto")
a = p.parse_args()

dev = torch.device("cuda", 0)
x = synth.gaussian_mixture_torch(a.rows, a.d, a.k, 1003, dev)
torch.manual_seed(1003)
km = KMeans(types.SimpleNamespace(computation=types.SimpleNamespace(device="cuda", num_gpus=1)), a.d, a.k,
            assign_mode=a.mode, warmup_rng="cuda")
km.to(dev)
km.lr = 1e-2
nb = a.rows // a.b
for i in range(a.settle):
    km.add(x[(i % nb) * a.b:(i % nb + 1) * a.b], sync=False)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
torch.cuda.profiler.start()
ev[0].record()
for i in range(a.settle, a.settle + a.steps):
    km.add(x[(i % nb) * a.b:(i % nb + 1) * a.b], sync=False)
ev[1].record()
for i in range(a.steps):
    km.calc_best(x[i * a.b:(i + 1) * a.b], sync=False)
ev[2].record()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("add ms/step %.3f  calc_best ms/batch %.3f  fallbacks %d  centroids in use %d  counts max %.0f" % (
    ev[0].elapsed_time(ev[1]) / a.steps, ev[1].elapsed_time(ev[2]) / a.steps, km.fallback,
    int((km.counts > 0).sum()), km.counts.max().item()))
hist = torch.bincount(km.last_best, minlength=a.k)
print("last batch: max rows/centroid %d, empty centroids %d" % (hist.max().item(), (hist == 0).sum().item()))
