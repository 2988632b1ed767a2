"""Per-CTA phase timers of one greedy-MI iteration of a stream loop, with the SM each CTA ran on:
    python tools/mi_cta_dump.py [loop] [variant] [w] [k]   -> one line per CTA, sorted by scan time"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from acav100m_b200 import _lib, synth
from acav100m_b200.subset_selection import get_measure

loop = sys.argv[1] if len(sys.argv) > 1 else "bytes"
variant = int(sys.argv[2]) if len(sys.argv) > 2 else 3
w = int(sys.argv[3]) if len(sys.argv) > 3 else 100_000_000
k = int(sys.argv[4]) if len(sys.argv) > 4 else 1024
cells = synth.zipf_pairs_torch(w, k, 1004, torch.device("cuda", 0))
sm = torch.cuda.get_device_properties(0).multi_processor_count
m = get_measure("mem_mi")(cells, ncentroids=k, device="cuda", loop=loop)
m.init_from_cells([(0, 1)], cells)
if loop == "bytes":
    _lib.call("acav_mi_set_stream_variant", m._engine, variant, 1)
m.select(40)
buf = torch.zeros(8 * sm + 64 * sm, dtype=torch.int64, device="cuda")
_lib.call("acav_mi_debug_timers", m._engine, _lib.ptr(buf))
m.select(8)
torch.cuda.synchronize()
allraw = buf.cpu().numpy()
raw = allraw[:8 * sm].reshape(sm, 8)
warps = allraw[8 * sm:].reshape(sm, 2, 32) / 1.965e3
us = raw[:, :4] / 1.965e3
order = np.argsort(us[:, 1])
print("cta smid rows blocks gain scan reduce barrier pre learn")
for c in order:
    print(c, int(raw[c, 5] >> 16), int(raw[c, 5] & 0xFFFF), int(raw[c, 4]), *[round(float(x), 1) for x in us[c]],
          round(raw[c, 6] / 1.965e3, 2), round(raw[c, 7] / 1.965e3, 2))

if loop == "bytes":
    print("per-warp (scan end, settled) in us since the iteration began, three CTAs:")
    for c in (order[0], order[len(order) // 2], order[-1]):
        print("cta", c, "scan end", [round(float(v), 1) for v in warps[c, 0]])
        print("cta", c, "settled ", [round(float(v), 1) for v in warps[c, 1]])
