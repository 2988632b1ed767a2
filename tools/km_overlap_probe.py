"""Do the background preparation kernel and the distance GEMM really share the SMs?  Runs each alone, then both
on two streams with no dependencies between them, sampling SM clock and board power:  python tools/km_overlap_probe.py [variant]"""
import sys
import torch
sys.path.insert(0, ".")
from acav100m_b200 import _lib, synth
from acav100m_b200.clustering import KMeans

variant = int(sys.argv[1]) if len(sys.argv) > 1 else 0
prep_fn = "acav_kmeans_prepare_batch"
n, d, k, chunk, reps = 262144, 2048, 1024, 131072, 10
dev = torch.device("cuda", 0)
x = synth.gaussian_mixture_torch(n, d, k, 1003, dev)
g = torch.Generator(device=dev).manual_seed(1003)
means = torch.randn(k, d, generator=g, device=dev) * 3.0
kms = []
for _ in range(2):
    km = KMeans(None, d, k, assign_mode="tensor", tile_variant=variant)
    km.to(dev)
    km.centers.copy_(means); km.counts.fill_(1000.0); km.count = 1000 * k
    kms.append(km)
ws = [km._workspace(chunk) for km in kms]
thr, r = kms[0].underused_threshold(), float(kms[0].reinit[1])
main = torch.cuda.current_stream(dev)
hot = torch.cuda.Stream(device=dev, priority=-1)
mp, hp = _lib.ctypes.c_void_p(main.cuda_stream), _lib.ctypes.c_void_p(hot.cuda_stream)
best = torch.empty(chunk, dtype=torch.int64, device=dev)
x0, x1 = x[:chunk], x[chunk:]
for h in ws:
    _lib.call("acav_kmeans_prepare_centers", h, _lib.ptr(means), _lib.ptr(kms[0].counts), thr, r, mp)
_lib.call("acav_kmeans_prepare_batch", ws[0], _lib.ptr(x0), chunk, d, mp)
torch.cuda.synchronize()


def gemm(sp):
    for _ in range(reps):
        _lib.call("acav_kmeans_assign_prepared", ws[0], _lib.ptr(x0), chunk, d, _lib.ptr(means), _lib.ptr(kms[0].counts),
                  thr, r, _lib.ptr(best), None, None, None, sp)


def prep(sp):
    for _ in range(reps):
        _lib.call(prep_fn, ws[1], _lib.ptr(x1), chunk, d, sp)


def run(which):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(main)
    hot.wait_event(e0)
    if which in ("gemm", "both"):
        gemm(hp)
    if which in ("prep", "both"):
        prep(mp)
    main.wait_stream(hot)
    e1.record(main)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


import subprocess
for w in ("gemm", "prep", "both"):
    run(w)
    reps = 3000
    smi = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "100"],
                           stdout=subprocess.PIPE, text=True)
    ms = run(w)
    smi.terminate()
    lines = [l.strip().split(",") for l in smi.stdout.read().strip().splitlines()]
    vals = [(float(a), float(b)) for a, b in lines if a.strip().replace(".", "").isdigit()]
    mid = vals[len(vals) // 3:] or vals
    clk = sorted(v[0] for v in mid)[len(mid) // 2]
    pw = sorted(v[1] for v in mid)[len(mid) // 2]
    print(f"variant {variant} {prep_fn[12:]}: {w:5s} {ms:.3f} ms per chunk of {chunk}; median SM clock {clk:.0f} MHz, "
          f"power {pw:.0f} W ({len(vals)} samples)", flush=True)
    reps = 10
