"""Check and time the three tcgen05 distance-GEMM tile shapes (acav_kmeans_set_tile_variant) on prepared operands:
    python tools/km_tile_bench.py [d] [k] [variants e.g. 1,2,3] [batches e.g. 8192,131072]
Times acav_kmeans_assign_prepared (GEMM + merge/classify + re-check of near-ties) with CUDA events."""
import sys
import torch
sys.path.insert(0, ".")
from acav100m_b200 import _lib
from acav100m_b200.clustering import KMeans

d = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
k = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
variants = [int(v) for v in (sys.argv[3] if len(sys.argv) > 3 else "1,2,3").split(",")]
batches = [int(v) for v in (sys.argv[4] if len(sys.argv) > 4 else "8192,131072").split(",")]
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)
means = torch.randn(k, d, generator=g, device=dev) * 3
names = {1: "single 128x256", 2: "pair 256x256", 3: "pair 256x512"}
for b in batches:
    x = means[torch.randint(0, k, (b,), generator=g, device=dev)] + torch.randn(b, d, generator=g, device=dev)
    ref = None
    for v in variants:
        km = KMeans(None, d, k, assign_mode="tensor", tile_variant=v)
        km.to(dev)
        km.centers.copy_(means); km.counts.fill_(1000.0); km.counts[::3] = 0.0; km.count = 1000 * k
        ws = km._workspace(b)
        thr, r = km.underused_threshold(), float(km.reinit[1])
        st = _lib.stream_ptr(dev)
        best = torch.empty(b, dtype=torch.int64, device=dev)
        nref = torch.zeros(2, dtype=torch.int32, device=dev)
        _lib.call("acav_kmeans_prepare_centers", ws, _lib.ptr(km.centers), _lib.ptr(km.counts), thr, r, st)
        _lib.call("acav_kmeans_prepare_batch", ws, _lib.ptr(x), b, d, st)

        def run():
            _lib.call("acav_kmeans_assign_prepared", ws, _lib.ptr(x), b, d, _lib.ptr(km.centers), _lib.ptr(km.counts),
                      thr, r, _lib.ptr(best), None, None, _lib.ptr(nref), st)
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 20
        e0.record()
        for _ in range(n):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        if ref is None:
            kme = KMeans(None, d, k, assign_mode="exact")
            kme.to(dev)
            kme.centers.copy_(km.centers); kme.counts.copy_(km.counts); kme.count = km.count
            m = min(b, 16384)
            ref = kme.calc_best(x[:m], distance=False)[0]
        m = ref.numel()
        ndiff = int((best[:m] != ref).sum())
        print(f"b={b:7d} d={d} k={k} variant {v} ({names[v]}): {ms*1e3:9.1f} us  {2.0*b*k*d/ms/1e9:8.1f} TFLOP/s  "
              f"refined {nref.tolist()}  rows differing from exact (first {m}): {ndiff}", flush=True)
        del km
