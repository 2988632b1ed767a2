"""Time the tcgen05 assignment kernel alone (through acav_kmeans_assign, converged-state data) for several batch
sizes: python tools/km_assign_sweep.py [d] [k]"""
import sys
import torch
sys.path.insert(0, ".")
from acav100m_b200.clustering import KMeans
d = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
k = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)
means = torch.randn(k, d, generator=g, device=dev) * 3
for b in (8192, 18944, 65536, 131072, 262144):
    x = means[torch.randint(0, k, (b,), generator=g, device=dev)] + torch.randn(b, d, generator=g, device=dev)
    km = KMeans(None, d, k, assign_mode="tensor")
    km.to(dev)
    km.centers.copy_(means); km.counts.fill_(1000.0); km.count = 1000 * k
    for _ in range(3):
        km.calc_best(x, sync=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    e0.record()
    for _ in range(n):
        km.calc_best(x, sync=False)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"b={b:7d} d={d} k={k}: calc_best {ms*1e3:9.1f} us  {2.0*b*k*d/ms/1e9:8.1f} TFLOP/s (whole call)", flush=True)
    del km, x
