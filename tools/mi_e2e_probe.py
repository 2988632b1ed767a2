"""Where the end-to-end greedy-MI time goes (bench.py `e2e`): pinned host int64 [W, 2] -> engine -> 20 picks -> host.
    python tools/mi_e2e_probe.py [w] [loop] [repeats]"""
import sys
import time

import torch

sys.path.insert(0, ".")
from acav100m_b200 import _lib, synth
from acav100m_b200.subset_selection import get_measure

w = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
loop = sys.argv[2] if len(sys.argv) > 2 else "bytes"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
k = 1024
dev = torch.device("cuda", 0)
cells = synth.zipf_pairs_torch(w, k, 1004, dev)
host = torch.empty((w, 2), dtype=torch.int64, pin_memory=True)
host.copy_(cells)
del cells
torch.cuda.synchronize()
for r in range(reps):
    t = [time.perf_counter()]
    m = get_measure("mem_mi")(host, ncentroids=k, device="cuda", loop=loop)
    m.init_from_cells([(0, 1)], host)
    torch.cuda.synchronize(); t.append(time.perf_counter())
    _lib.call("acav_mi_prepare", m._engine, m._loop_mode(), _lib.stream_ptr(dev))
    torch.cuda.synchronize(); t.append(time.perf_counter())
    pos, gain = m.select(20)
    torch.cuda.synchronize(); t.append(time.perf_counter())
    ph, gh = pos.cpu(), gain.cpu()
    t.append(time.perf_counter())
    print("rep %d: H2D+pack+tables %.1f ms, layout build %.1f ms, 20 picks %.2f ms, D2H %.2f ms, total %.1f ms"
          % (r, *(1e3 * (t[i + 1] - t[i]) for i in range(4)), 1e3 * (t[-1] - t[0])), flush=True)
    del m
    torch.cuda.synchronize()
