"""Time the multi-pair greedy-MI engine (acav_mi_pairs_*): python tools/mi_pairs_bench.py [W D C iters]
Prints one JSON line: microseconds per greedy iteration and the bytes / gathers one iteration moves."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, __file__.rsplit("/tools/", 1)[0])
from acav100m_b200.subset_selection import get_measure          # noqa: E402
from acav100m_b200.subset_selection.pairing import get_cluster_pairing   # noqa: E402


def main():
    W, D, C, iters = (int(x) for x in (sys.argv[1:5] + ["1000000", "10", "256", "200"][len(sys.argv) - 1:]))
    rng = np.random.RandomState(1)
    base = rng.randint(0, C, size=(W, 1))
    a = np.where(rng.random_sample((W, D)) < 0.5, (base + np.arange(D)) % C, rng.randint(0, C, size=(W, D))).astype(np.int64)
    pairs = get_cluster_pairing([("m%d" % i, "layer") for i in range(D)], "combination")
    m = get_measure("mem_mi")(torch.from_numpy(a), ncentroids=C, device="cuda")
    m.init(pairs, torch.arange(W))
    m.select(20)                                                  # warm-up
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    pos, gain = m.select(iters)
    t1.record()
    torch.cuda.synchronize()
    us = 1e3 * t0.elapsed_time(t1) / iters
    P = len(pairs)
    print(json.dumps({"W": W, "D": D, "C": C, "P": P, "iters": iters, "us_per_iteration": us,
                      "candidate_pairs_per_sec": W * P / (us * 1e-6), "candidates_per_sec": W / (us * 1e-6),
                      "id_stream_bytes_per_iteration": 2 * D * W, "score_table_bytes": 4 * P * C * C,
                      "first_picks": pos[:5].tolist(), "last_gain": float(gain[-1])}))


if __name__ == "__main__":
    main()
