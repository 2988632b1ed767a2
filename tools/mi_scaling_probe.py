"""Microseconds per greedy-MI iteration for a list of per-GPU sizes, one GPU or sharded under torchrun, with the per-CTA
phase timers of the stream loops (rank 0):
    python tools/mi_scaling_probe.py 12500000,50000000,100000000
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 tools/mi_scaling_probe.py 12500000,50000000"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from acav100m_b200 import _lib, synth
from acav100m_b200.subset_selection import get_measure

sizes = [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "12500000,50000000,100000000").split(",")]
loops = (sys.argv[2] if len(sys.argv) > 2 else "bytes,persistent,cells").split(",")
k, picks, warm = 1024, 100, int(os.environ.get("PREWARM", 20))
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
dist = None
if world > 1:
    import torch.distributed as dist
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
dev = torch.device("cuda", torch.cuda.current_device())
sm = torch.cuda.get_device_properties(dev).multi_processor_count
for w in sizes:
    cells = synth.zipf_pairs_torch(w, k, 1004 + rank, dev)
    for loop in loops:
        m = get_measure("mem_mi")(cells, ncentroids=k, device="cuda", shard=(rank, world) if world > 1 else None, loop=loop)
        m.init_from_cells([(0, 1)], cells, w_global=w * world, lo=w * rank)
        m.select(warm)
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        m.select(picks)
        e1.record()
        torch.cuda.synchronize()
        line = {"world": world, "w_per_gpu": w, "loop": m.loop_name(), "us_per_iteration": round(e0.elapsed_time(e1) * 1e3 / picks, 2)}
        if loop != "cells":
            buf = torch.zeros(72 * sm, dtype=torch.int64, device=dev)
            _lib.call("acav_mi_debug_timers", m._engine, _lib.ptr(buf))
            m.select(8)
            torch.cuda.synchronize()
            _lib.call("acav_mi_debug_timers", m._engine, None)
            raw = buf.cpu().numpy()[:8 * sm].reshape(sm, 8).astype(np.float64)
            t = raw[:, :4] / 1.965e3
            for j, n in enumerate(("gain rows", "scan", "reduce+publish", "barrier+exchange wait")):
                line[n] = [round(float(t[:, j].min()), 1), round(float(t[:, j].mean()), 1), round(float(t[:, j].max()), 1)]
            line["prologue"] = round(float(raw[:, 6].mean() / 1.965e3), 2)
            line["learn"] = round(float(raw[:, 7].mean() / 1.965e3), 2)
        if rank == 0:
            print(json.dumps(line), flush=True)
        del m
    del cells
if dist:
    dist.barrier()
    dist.destroy_process_group()
