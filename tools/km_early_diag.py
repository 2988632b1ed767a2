"""Why do rows of the early training state need the exact re-check?  Trains from the reference init, then for one batch
prints how the screen classified the rows and how many centroids lie inside the bf16 error bound of the best one.
    python tools/km_early_diag.py [settle steps] [b] [d] [k]"""
import sys, types
import torch
sys.path.insert(0, ".")
from acav100m_b200 import _lib, synth
from acav100m_b200.clustering import KMeans
settle = int(sys.argv[1]) if len(sys.argv) > 1 else 60
b = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
d = int(sys.argv[3]) if len(sys.argv) > 3 else 2048
k = int(sys.argv[4]) if len(sys.argv) > 4 else 1024
dev = torch.device("cuda", 0)
x = synth.gaussian_mixture_torch(400_000, d, k, 1003, dev)
torch.manual_seed(1003)
km = KMeans(types.SimpleNamespace(computation=types.SimpleNamespace(device="cuda", num_gpus=1)), d, k,
            assign_mode="tensor", warmup_rng="cuda")
km.to(dev); km.lr = 1e-2
nb = x.shape[0] // b
for s in range(settle + 1):
    if s in (2, 5, 10, 20, 40, settle):
        xb = x[(s % nb) * b:(s % nb + 1) * b]
        ws = km._workspace(b)
        best = torch.empty(b, dtype=torch.int64, device=dev)
        nref = torch.zeros(2, dtype=torch.int32, device=dev)
        thr = km.underused_threshold()
        _lib.call("acav_kmeans_assign", ws, _lib.ptr(xb), b, d, _lib.ptr(km.centers), _lib.ptr(km.counts), thr,
                  float(km.reinit[1]), _lib.ptr(best), None, None, _lib.ptr(nref), _lib.ASSIGN_TENSOR, _lib.stream_ptr(dev))
        scale = torch.where(km.counts < thr, 1.0 / km.reinit[1], 1.0)
        cn = (km.centers.double() ** 2).sum(1)
        xn = (xb[:2048].double() ** 2).sum(1)
        dist = (-2.0 * xb[:2048].double() @ km.centers.double().T + xn[:, None] + cn[None, :]) * scale[None, :].double()
        srt, _ = dist.sort(dim=1)
        cmax = cn.max().sqrt()
        bound = 5.0 / 256.0 * xn.sqrt() * cmax
        within = (dist <= (srt[:, :1] + bound[:, None])).sum(1).float()
        own = 2.5 / 256.0 * 2 * xn.sqrt()[:, None] * cn.sqrt()[None, :] * scale[None, :].double()     # per-centroid bound
        within_pc = (dist - own <= (srt[:, :1] + own.gather(1, dist.argmin(1, keepdim=True)))).sum(1).float()
        print(f"step {s:3d}: count={km.count} underused={(km.counts < thr).sum().item():4d} |c| min/med/max "
              f"{cn.sqrt().min():.3g}/{cn.sqrt().median():.3g}/{cn.sqrt().max():.3g}  refined(cand,full)={nref.tolist()} of {b}; "
              f"gap d2-d1 median {float((srt[:,1]-srt[:,0]).median()):.4g} bound median {float(bound.median()):.4g}; "
              f"centroids within global bound: median {within.median():.0f} max {within.max():.0f}; "
              f"with per-centroid bound: median {within_pc.median():.0f} max {within_pc.max():.0f}", flush=True)
    km.add(x[(s % nb) * b:(s % nb + 1) * b], sync=False)
