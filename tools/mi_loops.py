"""Time the greedy-MI loops against each other on one candidate list and check that they pick the same clips:
    python tools/mi_loops.py [w] [k] [picks] [warm]"""
import sys
import torch
sys.path.insert(0, ".")
from acav100m_b200 import synth
from acav100m_b200.subset_selection import get_measure
w = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
picks = int(sys.argv[3]) if len(sys.argv) > 3 else 2000
warm = int(sys.argv[4]) if len(sys.argv) > 4 else 20
cells = synth.zipf_pairs_torch(w, k, 1004, torch.device("cuda", 0))
out = {}
for loop in ("persistent", "cells"):
    m = get_measure("mem_mi")(cells, ncentroids=k, device="cuda", loop=loop)
    m.init_from_cells([(0, 1)], cells)
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True); t2 = torch.cuda.Event(enable_timing=True)
    t0.record()
    p0, g0 = m.select(warm)          # includes building the stream / the cell index
    t1.record()
    p1, g1 = m.select(picks)
    t2.record()
    torch.cuda.synchronize()
    out[loop] = (torch.cat([p0, p1]), torch.cat([g0, g1]))
    print(f"{loop:10s}: build + {warm} picks {t0.elapsed_time(t1):9.2f} ms; {picks} picks {t1.elapsed_time(t2):9.2f} ms "
          f"= {t1.elapsed_time(t2) * 1e3 / picks:7.2f} us per iteration", flush=True)
    del m
same = torch.equal(out["persistent"][0], out["cells"][0]) and torch.equal(out["persistent"][1], out["cells"][1])
print("identical picks and gains:", same)
