"""Where does the assignment pass go?  Times, over a resident [n, d] fp32 shard: the fp32->bf16 preparation alone,
the prepared distance GEMM + classify alone, and KMeans.assign_all (both, overlapped on two streams).
    python tools/km_pass_parts.py [n] [d] [k] [variant] [chunk]"""
import sys
import torch
sys.path.insert(0, ".")
from acav100m_b200 import _lib, synth
from acav100m_b200.clustering import KMeans

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_250_000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
k = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
variant = int(sys.argv[4]) if len(sys.argv) > 4 else 0
chunk = int(sys.argv[5]) if len(sys.argv) > 5 else 131072
dev = torch.device("cuda", 0)
x = synth.gaussian_mixture_torch(n, d, k, 1003, dev)
g = torch.Generator(device=dev).manual_seed(1003)
means = torch.randn(k, d, generator=g, device=dev) * 3.0
km = KMeans(None, d, k, assign_mode="tensor", tile_variant=variant)
km.to(dev)
km.centers.copy_(means); km.counts.fill_(1000.0); km.count = 1000 * k
ws = km._workspace(chunk)
thr, r = km.underused_threshold(), float(km.reinit[1])
st = _lib.stream_ptr(dev)
best = torch.empty(n, dtype=torch.int64, device=dev)


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def prep_all():
    for lo in range(0, n, chunk):
        xb = x[lo:lo + chunk]
        _lib.call("acav_kmeans_prepare_batch", ws, _lib.ptr(xb), xb.shape[0], d, st)


def gemm_all():
    for lo in range(0, n, chunk):
        xb = x[lo:lo + chunk]
        _lib.call("acav_kmeans_assign_prepared", ws, _lib.ptr(xb), xb.shape[0], d, _lib.ptr(km.centers),
                  _lib.ptr(km.counts), thr, r, _lib.c_vp(best.data_ptr() + 8 * lo), None, None, None, st)


_lib.call("acav_kmeans_prepare_centers", ws, _lib.ptr(km.centers), _lib.ptr(km.counts), thr, r, st)
flop = 2.0 * n * k * d
t_prep = timed(prep_all)
t_gemm = timed(gemm_all)
t_pass = timed(lambda: km.assign_all(x, chunk=chunk))
print(f"n={n} d={d} k={k} variant={variant} chunk={chunk}: prep {t_prep:.3f} ms ({n*d*6/t_prep/1e6:.0f} GB/s)  "
      f"gemm+classify {t_gemm:.3f} ms ({flop/t_gemm/1e9:.0f} TFLOP/s)  assign_all {t_pass:.3f} ms "
      f"({flop/t_pass/1e9:.0f} TFLOP/s)", flush=True)
