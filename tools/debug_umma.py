"""Diagnostic for the tcgen05 assignment kernel (not a test): compares tensor-mode assignment with the
exact mode and with a torch evaluation of the bf16-rounded distances."""
import sys
import torch

sys.path.insert(0, ".")
from acav100m_b200 import _lib
from acav100m_b200.clustering import KMeans


def run(b, d, k, seed=0, spread=3.0, clustered=True):
    torch.manual_seed(seed)
    if clustered:                        # centroids near the component means (a trained model)
        means = torch.randn(k, d, device="cuda") * spread
        x = means[torch.randint(0, k, (b,), device="cuda")] + torch.randn(b, d, device="cuda")
        c = means + 0.05 * torch.randn(k, d, device="cuda")
    else:
        x = torch.randn(b, d, device="cuda"); c = torch.randn(k, d, device="cuda")
    km = KMeans(None, d, k)
    km.to("cuda")
    km.centers.copy_(c)
    km.counts.fill_(100.0)
    km.counts[::3] = 0.0
    km.count = 50 * k
    outs = {}
    for mode in ("exact", "tensor"):
        km.assign_mode = mode
        ws = km._workspace(b)
        best = torch.empty(b, dtype=torch.int64, device="cuda")
        mind = torch.empty(b, dtype=torch.float32, device="cuda")
        mean = torch.empty(1, dtype=torch.float32, device="cuda")
        nref = torch.zeros(2, dtype=torch.int32, device="cuda")
        _lib.call("acav_kmeans_assign", ws, _lib.ptr(x), b, d, _lib.ptr(km.centers), _lib.ptr(km.counts),
                  km.underused_threshold(), 5.0, _lib.ptr(best), _lib.ptr(mind), _lib.ptr(mean), _lib.ptr(nref),
                  km._mode(), _lib.stream_ptr())
        torch.cuda.synchronize()
        outs[mode] = (best.cpu(), mind.cpu(), mean.item(), nref.cpu().tolist())
    be, me, mne, _ = outs["exact"]
    bt, mt, mnt, nref = outs["tensor"]
    # torch view of the screen: bf16 inputs, fp32 math
    xb, cb = x.bfloat16().float(), c.bfloat16().float()
    dist = -2 * cb @ xb.T + (x * x).sum(1)[None] + (c * c).sum(1)[:, None]
    under = km.counts < km.underused_threshold()
    dist[under] /= 5
    bscreen = dist.argmin(0).cpu()
    print(f"b={b} d={d} k={k} clustered={clustered}: ids tensor==exact {(be == bt).float().mean():.6f} "
          f"screen(torch bf16)==exact {(bscreen == be).float().mean():.6f} refined cand/full {nref} ({nref[0] / b:.4f}/{nref[1] / b:.4f}) "
          f"mean exact {mne:.6f} tensor {mnt:.6f} max|mind diff| {(me - mt).abs().max():.3e}")
    return bool((be == bt).all())


if __name__ == "__main__":
    ok = True
    for args in [(128, 64, 16), (256, 64, 256), (1000, 128, 256), (300, 88, 13), (4097, 512, 300),
                 (8192, 2048, 1024), (20000, 2048, 1024)]:
        ok &= run(*args)
    ok &= run(4096, 256, 512, clustered=False)
    print("ALL OK" if ok else "MISMATCH")
