"""Time every greedy-MI loop / stream variant on one candidate list in one process, check that all pick the same clips,
and print the per-CTA phase timers of the stream loops:
    python tools/mi_variants.py [w] [k] [picks] [warm]  > gpurun_out/mi_variants.jsonl"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from acav100m_b200 import _lib, synth
from acav100m_b200.subset_selection import get_measure

w = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
picks = int(sys.argv[3]) if len(sys.argv) > 3 else 200
warm = int(sys.argv[4]) if len(sys.argv) > 4 else 20
cells = synth.zipf_pairs_torch(w, k, 1004, torch.device("cuda", 0))
sm = torch.cuda.get_device_properties(0).multi_processor_count
ref = None

CONFIGS = [("persistent", None, None, None), ("cells", None, None, None)]
for variant in (0, 1, 2, 3, 4, 5):
    CONFIGS.append(("bytes", variant, 1, None))



def timers(m, names=("gain rows", "scan", "reduce+publish", "barrier wait")):
    buf = torch.zeros(72 * sm, dtype=torch.int64, device="cuda")
    _lib.call("acav_mi_debug_timers", m._engine, _lib.ptr(buf))
    m.select(8)
    torch.cuda.synchronize()
    _lib.call("acav_mi_debug_timers", m._engine, None)
    raw = buf.cpu().numpy()[:8 * sm].reshape(sm, 8).astype(np.float64)
    t = raw[:, :4] / 1.965e3
    out = {n: {"min": round(float(t[:, j].min()), 1), "mean": round(float(t[:, j].mean()), 1), "max": round(float(t[:, j].max()), 1)}
           for j, n in enumerate(names)}
    out["prologue_us_mean"] = round(float(raw[:, 6].mean() / 1.965e3), 2)
    out["learn_us_mean"] = round(float(raw[:, 7].mean() / 1.965e3), 2)
    out["rows_per_cta_max"] = int(raw[:, 5].max())
    out["blocks_per_cta_min_max"] = [int(raw[:, 4].min()), int(raw[:, 4].max())]
    return out


for loop, variant, cache, rowcost in CONFIGS:
    if rowcost is not None:
        os.environ["ACAV_MI_S8_ROWCOST"] = str(rowcost)
    else:
        os.environ.pop("ACAV_MI_S8_ROWCOST", None)
    try:
        m = get_measure("mem_mi")(cells, ncentroids=k, device="cuda", loop=loop)
        m.init_from_cells([(0, 1)], cells)
        if variant is not None:
            _lib.call("acav_mi_set_stream_variant", m._engine, variant, cache)
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        p0, g0 = m.select(warm)
        e[1].record()
        p1, g1 = m.select(picks)
        e[2].record()
        torch.cuda.synchronize()
        got = (torch.cat([p0, p1]), torch.cat([g0, g1]))
        if ref is None:
            ref = got
        line = {"loop": loop, "variant": variant, "cache": cache, "rowcost": rowcost,
                "us_per_iteration": round(e[1].elapsed_time(e[2]) * 1e3 / picks, 2),
                "build_plus_warm_ms": round(e[0].elapsed_time(e[1]), 2),
                "same_picks_and_gains_as_first": bool(torch.equal(got[0], ref[0]) and torch.equal(got[1], ref[1]))}
        if loop != "cells":
            line["timers_us"] = timers(m)
        del m
    except Exception as ex:                                   # report, keep sweeping
        line = {"loop": loop, "variant": variant, "cache": cache, "rowcost": rowcost, "error": repr(ex)[:300]}
    print(json.dumps(line), flush=True)
