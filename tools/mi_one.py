"""One short greedy-MI selection for ncu captures: python tools/mi_one.py [w] [k] [picks] [loop] [warm]"""
import sys
import torch
sys.path.insert(0, ".")
from acav100m_b200 import synth
from acav100m_b200.subset_selection import get_measure
w = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
picks = int(sys.argv[3]) if len(sys.argv) > 3 else 4
loop = sys.argv[4] if len(sys.argv) > 4 else "persistent"
warm = int(sys.argv[5]) if len(sys.argv) > 5 else picks
cells = synth.zipf_pairs_torch(w, k, 1004, torch.device("cuda", 0))
m = get_measure("mem_mi")(cells, ncentroids=k, device="cuda", loop=loop)
m.init_from_cells([(0, 1)], cells)
m.select(warm)
torch.cuda.synchronize()
torch.cuda.profiler.start()
m.select(picks)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
