"""BASELINE config 2 end to end on one B200, synthetic features resident in HBM:

    python tools/run_config2.py [--clips 1000000] [--k 256] [--select 100000] [--batch 65536]

k-means (K = 256) on the audio (D = 512) and visual (D = 2048) features -- one epoch of SGD steps at a large batch, then the
assignment pass -- followed by greedy-MI selection of `--select` clips from the resulting (audio id, visual id) pairs.
Prints one JSON line with the time of every stage.  A measurement aid, not a test: nothing is compared here (parity lives
in tests/); the stages are the public operators (`KMeans.add`, `KMeans.assign_all`, `get_measure('mem_mi')`).
"""
import argparse
import json
import os
import sys
import time
import types

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from acav100m_b200 import synth                                   # noqa: E402
from acav100m_b200.clustering import KMeans                       # noqa: E402
from acav100m_b200.subset_selection import get_measure            # noqa: E402


def stage(fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    return out, time.perf_counter() - t0


def cluster(n, d, k, batch, seed, dev):
    x = synth.gaussian_mixture_torch(n, d, k, seed, dev)
    kargs = types.SimpleNamespace(computation=types.SimpleNamespace(device="cuda", num_gpus=1))
    torch.manual_seed(seed)
    km = KMeans(kargs, d, k, warmup_rng="cuda")
    km.to(dev)
    km.lr = 1e-2                                                   # epoch 0 of run_clustering.py:168

    def train():
        for lo in range(0, n - batch + 1, batch):
            km.add(x[lo:lo + batch], sync=False, distance=False)

    _, t_train = stage(train)
    ids, t_assign = stage(lambda: km.assign_all(x))
    del x
    return ids, {"d": d, "train_s": t_train, "steps": n // batch, "assign_s": t_assign,
                 "assign_tflops": 2.0 * n * k * d / t_assign / 1e12, "lr_fallbacks": km.fallback,
                 "clusters_used": int(torch.unique(ids).numel())}


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--clips", type=int, default=1_000_000)
    p.add_argument("--k", type=int, default=256)
    p.add_argument("--select", type=int, default=100_000)
    p.add_argument("--batch", type=int, default=65_536)
    p.add_argument("--loop", default="auto")
    args = p.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    ida, audio = cluster(args.clips, 512, args.k, args.batch, 2001, dev)
    idv, visual = cluster(args.clips, 2048, args.k, args.batch, 2002, dev)
    cells = torch.stack([ida, idv], dim=1).contiguous()
    # run_greedy.py:43-53: the first clip seeds S and is not a candidate; subset_size - 2 greedy picks follow
    m = get_measure("mem_mi")(cells, ncentroids=args.k, device="cuda", loop=args.loop)
    _, t_build = stage(lambda: m.init_from_cells([(0, 1)], cells[1:], max_picks=args.select + 8))
    n_picks = args.select - 2
    (pos, gain), t_select = stage(lambda: m.select(n_picks))
    scored = n_picks * (args.clips - 1) - n_picks * (n_picks - 1) // 2
    print(json.dumps({
        "config": "C2: %d clips, D_a=512, D_v=2048, K=%d, select %d, one B200" % (args.clips, args.k, args.select),
        "kmeans_audio": audio, "kmeans_visual": visual,
        "mi": {"loop": m.loop_name(), "build_s": t_build, "select_s": t_select, "picks": n_picks,
               "us_per_iteration": 1e6 * t_select / n_picks, "candidate_clips_per_sec": scored / t_select,
               "last_gain": float(gain[-1]), "first_picks": pos[:5].tolist()},
        "total_s": audio["train_s"] + audio["assign_s"] + visual["train_s"] + visual["assign_s"] + t_build + t_select,
    }))


if __name__ == "__main__":
    main()
