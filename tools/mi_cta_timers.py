"""Per-CTA phase timers of the persistent greedy-MI loop (load balance): python tools/mi_cta_timers.py [w] [k] [warm]"""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from acav100m_b200 import _lib, synth
from acav100m_b200.subset_selection import get_measure
w = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
warm = int(sys.argv[3]) if len(sys.argv) > 3 else 2000
cells = synth.zipf_pairs_torch(w, k, 1004, torch.device("cuda", 0))
m = get_measure("mem_mi")(cells, ncentroids=k, device="cuda", loop="persistent")
m.init_from_cells([(0, 1)], cells)
m.select(warm)
sm = torch.cuda.get_device_properties(0).multi_processor_count
buf = torch.zeros(72 * sm, dtype=torch.int64, device="cuda")
_lib.call("acav_mi_debug_timers", m._engine, _lib.ptr(buf))
m.select(8)
torch.cuda.synchronize()
raw = buf.cpu().numpy()[:8 * sm].reshape(sm, 8).astype(np.float64)
t = raw[:, :4] / 1.965e3        # us at 1965 MHz
names = ["gain rows", "scan", "reduce+publish", "barrier wait"]
for j, n in enumerate(names):
    c = t[:, j]
    print(f"{n:16s} min {c.min():8.1f} mean {c.mean():8.1f} max {c.max():8.1f} us   argmax CTA {c.argmax()}")
busy = t[:, :3].sum(1)
order = np.argsort(-busy)[:8]
print("slowest CTAs (busy us):", [(int(i), round(float(busy[i]), 1), [round(float(x), 1) for x in t[i, :3]]) for i in order])
print("fastest CTAs (busy us):", [(int(i), round(float(busy[i]), 1)) for i in np.argsort(busy)[:5]])

print("prologue (col terms) us: mean %.2f max %.2f; learn-winner (previous iteration) us: mean %.2f max %.2f" % (
    raw[:, 6].mean() / 1.965e3, raw[:, 6].max() / 1.965e3, raw[:, 7].mean() / 1.965e3, raw[:, 7].max() / 1.965e3))
print("CTA: blocks rows | gain scan reduce wait (us)")
for i in list(range(0, sm, 12)) + list(range(sm - 10, sm)):
    print(f"{i:4d}: {int(raw[i,4]):6d} {int(raw[i,5]):4d} | " + " ".join(f"{x:6.1f}" for x in t[i]))
