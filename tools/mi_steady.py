"""Per-iteration time of the greedy-MI loops at several points of a selection (not a test).

    python tools/mi_steady.py [--w 100000000 --k 1024 --steps 200 --at 0,2000,20000 --loops persistent,kernels]
"""
import argparse
import sys

import torch

sys.path.insert(0, ".")
from acav100m_b200 import synth
from acav100m_b200.subset_selection import get_measure

p = argparse.ArgumentParser()
p.add_argument("--w", type=int, default=100_000_000)
p.add_argument("--k", type=int, default=1024)
p.add_argument("--steps", type=int, default=200)
p.add_argument("--at", default="0,2000,20000")
p.add_argument("--loops", default="persistent,kernels")
a = p.parse_args()
dev = torch.device("cuda", 0)
cells = synth.zipf_pairs_torch(a.w, a.k, 1004, dev)
for loop in a.loops.split(","):
    m = get_measure("mem_mi")(cells, ncentroids=a.k, device="cuda", loop=loop)
    m.init_from_cells([(0, 1)], cells)
    done = 0
    for at in [int(x) for x in a.at.split(",")]:
        if at > done:
            m.loop = "persistent"
            m.select(at - done)
            done = at
        m.loop = loop
        m.select(3)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        m.select(a.steps)
        e1.record()
        torch.cuda.synchronize()
        done += a.steps + 3
        us = e0.elapsed_time(e1) * 1e3 / a.steps
        print(f"{loop:10s} W={a.w} K={a.k} after {at:6d} picks: {us:8.1f} us/iter  "
              f"{a.w / us * 1e6 / 1e9:8.1f} G cand/s  stream {2 * a.w / us / 1e3:7.1f} GB/s (2 B/cand) "
              f"{4 * a.w / us / 1e3:7.1f} GB/s (4 B/cand)", flush=True)
    del m
