#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on B200: greedy-MI candidate-clips/s (+ k-means iter/s), K = 1024.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload default|c2|c3|c4]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config.workload, DESIGN.md "Measurement"):
  * greedy MI: BASELINE config 4's candidate list -- 100 M (c_a, c_v) cluster-id pairs, K_a = K_v =
    1024 -- PER GPU (weak scaling; larger than L2, so every iteration streams from HBM).  One step =
    one exact greedy iteration: score every remaining candidate, first arg-max, table update, removal.
    value = candidates scored per second over the K timed steps, summed over ranks, through the candidate-STREAM
    loop (the kernel north_star describes); the same iterations through the cell-index loop are reported next to it
    (`cell_index_loop`), and `c4` holds config 4 as written -- 1e8 candidates in TOTAL, sharded over the ranks (strong
    scaling) -- as microseconds per iteration for both loops.
  * k-means (reported under "kmeans", summarised in roofline.kmeans_* / e2e.kmeans so the driver's record keeps it):
    BASELINE config 3's per-GPU shard -- 1.25 M x 2048 fp32 rows resident in HBM, K = 1024, per-GPU batch 8192
    (global 65 536 at 8 GPUs) -- trained by KMeans.add itself from the reference's init for two epochs before
    anything is timed; SGD steps/s via KMeans.add and assignment rows/s via KMeans.calc_best / assign_all.
  * --workload c2 / c3 / c4 run BASELINE configs 2, 3 (one epoch + assignment pass) and 4 (1e8 -> 1e7 picks to
    completion) end to end instead and print their own line.
Inputs are synthetic (acav100m_b200.synth) and resident in HBM when the timed region starts; "e2e"
repeats the measurement through the public API starting from pinned HOST buffers (a whole job -- host list -> engine
-> `steps` picks -> results on the host -- timed once, after one identical untimed job).
`--impl reference` times the reference's CPU implementation (the oracle port; the reference is pure
Python and does not travel to the GPU box) on the host cores for the same metric.
"""
import argparse
import gc
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "greedy_mi_candidate_clips_per_sec"
UNIT = "candidate-clips/s"


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=200)
    p.add_argument("--warmup", type=int, default=20)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--workload", default="default", choices=["default", "c2", "c3", "c4"])
    p.add_argument("--k", type=int, default=1024)
    p.add_argument("--mi-candidates", type=int, default=100_000_000, help="per GPU")
    p.add_argument("--later-picks", type=int, default=2000, help="untimed picks before the `later_iterations` measurement (0 = skip)")
    p.add_argument("--mi-loop", default="bytes", help="headline loop: bytes (1-byte candidate stream) / persistent (2-byte stream) / cells / auto")
    p.add_argument("--km-rows", type=int, default=1_250_000, help="per GPU, resident")
    p.add_argument("--km-d", type=int, default=2048)
    p.add_argument("--km-batch", type=int, default=8192, help="per GPU")
    p.add_argument("--km-steps", type=int, default=0, help="0 = min(steps, rows/batch)")
    p.add_argument("--km-epochs", type=int, default=2, help="training epochs over the resident shard before timing")
    p.add_argument("--km-mode", default="auto")
    p.add_argument("--km-graph", default="auto", choices=["auto", "off"])
    p.add_argument("--km-comm", default="auto", choices=["auto", "p2p", "nccl"])
    p.add_argument("--cpu-sample", type=int, default=10_000_000)
    p.add_argument("--c4-picks", type=int, default=10_000_000)
    p.add_argument("--skip-cpu-baseline", action="store_true")
    p.add_argument("--skip-kmeans", action="store_true")
    p.add_argument("--skip-mi", action="store_true", help="development only: the line then has no headline value")
    p.add_argument("--skip-e2e", action="store_true")
    p.add_argument("--skip-cells", action="store_true")
    p.add_argument("--skip-strong", action="store_true")
    p.add_argument("--skip-parity", action="store_true")
    p.add_argument("--parity-candidates", type=int, default=2_000_000, help="per rank, for the parity_n self-check")
    p.add_argument("--parity-picks", type=int, default=256)
    return p.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained"), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def ncu_traffic(kernel):
    """DRAM bytes per launch of `kernel` from the ncu capture the profiling tool wrote (profiles/ncu_traffic.json,
    stamped with the commit it was taken at) -- never a constant in this file.  None when no capture is on record."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        return json.load(open(path)).get(kernel)
    except (OSError, ValueError):
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
            # nvidia-smi's start-up (NVML initialisation over every GPU of the box) takes the driver's locks for a few
            # hundred ms; a sub-millisecond timed region that falls into it pays for that.  Wait for the first sample.
            t_end = time.time() + 5.0
            while not self.lines and time.time() < t_end and self.proc.poll() is None:
                time.sleep(0.01)
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def dist_setup(n):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        return dist, rank, world, local
    if n > 1:
        raise SystemExit("--gpus %d needs torchrun (one rank per GPU)" % n)
    torch.cuda.set_device(0)
    return None, 0, 1, 0


def timed(dist, fn_warm, fn_timed):
    """barrier + sync, CUDA events around fn_timed on the current stream, max over ranks (ms)."""
    fn_warm()
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    out = fn_timed()
    e1.record()
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
    if dist:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item()), out


# -------------------------------------------------------------------------------------------------
# greedy MI
# -------------------------------------------------------------------------------------------------

def mi_engine(cells_src, k, rank, world, loop, w_global, lo, max_picks=None):
    from acav100m_b200.subset_selection import get_measure
    shard = (rank, world) if world > 1 else None
    m = get_measure("mem_mi")(cells_src, ncentroids=k, device="cuda", shard=shard, loop=loop)
    m.init_from_cells([(0, 1)], cells_src, w_global=w_global, lo=lo, max_picks=max_picks)
    return m


def mi_parity(args, dist, rank, world, timed_pos, timed_gain):
    """Self-certification of a (multi-GPU) run, printed as `parity_n` in the JSON line:
      1. every rank returned the same (positions, gains) from the timed iterations;
      2. a 2e6-candidates-per-rank list, sharded exactly like the timed one, run through every persistent loop, gives
         the picks and fp32 gains of the C oracle (oracle/mi_oracle.c, bucketed scan) on the WHOLE list, bit for bit.
    The oracle is the checker here, not the thing measured."""
    from acav100m_b200 import synth
    out = {"world": world}
    ok = True
    if world > 1:
        mine = torch.cat([timed_pos.to(torch.float64), timed_gain.to(torch.float64)])
        allr = torch.empty(world * mine.numel(), dtype=torch.float64, device=mine.device)
        dist.all_gather_into_tensor(allr, mine)
        same = bool((allr.view(world, -1) == allr.view(world, -1)[0]).all().item())
        out["timed_picks_identical_on_all_ranks"] = same
        ok &= same
    w_small, picks = args.parity_candidates, args.parity_picks
    lists = [synth.zipf_pairs(w_small, args.k, 7000 + r) for r in range(world)]      # every rank can rebuild every shard
    want = None
    if rank == 0:
        want = cpu_oracle_picks(np.concatenate(lists), args.k, picks)
    loops = {}
    for loop in ("persistent", "cells", "bytes"):
        cells = torch.from_numpy(lists[rank]).cuda()
        m = mi_engine(cells, args.k, rank, world, loop, w_small * world, w_small * rank)
        pos, gain = m.select(picks)
        m.check_status()
        if rank == 0:
            good = bool(np.array_equal(pos.cpu().numpy(), want[0]) and np.array_equal(gain.cpu().numpy(), want[1]))
            loops[m.loop_name()] = good
            ok &= good
        del m
    out["vs_c_oracle_bit_exact"] = loops
    out["what"] = ("%d picks from %d candidates per rank x %d ranks vs oracle/mi_oracle.c on the whole list"
                   % (picks, w_small, world))
    out["result"] = "ok" if ok else "MISMATCH"
    return out


def time_loop(args, dist, rank, world, cells, loop, w_per_rank):
    """(ms of `steps` iterations after `warmup`, their picks, engine) for one loop over this rank's first w_per_rank
    candidates."""
    m = mi_engine(cells[:w_per_rank], args.k, rank, world, loop, w_per_rank * world, w_per_rank * rank)
    ms, out = timed(dist, lambda: m.select(args.warmup), lambda: m.select(args.steps))
    m.check_status()
    return ms, out, m


def run_mi(args, dist, rank, world):
    from acav100m_b200 import synth
    W = args.mi_candidates
    dev = torch.device("cuda", torch.cuda.current_device())
    cells = synth.zipf_pairs_torch(W, args.k, 1004 + rank, dev)
    ms, (t_pos, t_gain), m = time_loop(args, dist, rank, world, cells, args.mi_loop, W)
    w_global = W * world
    scored = sum(w_global - args.warmup - i for i in range(args.steps))
    # algorithmic bytes per iteration and GPU in the layout the loop actually streams:
    # persistent = 2-byte c2 stream, bytes = 1-byte sub-row stream, list order = 4-byte packed pairs; + 4*K^2 table
    name = m.loop_name()
    bytes_per_cand = 2.0 if name.startswith("persistent") else 1.0 if name.startswith("bytes") else 4.0
    per_iter_bytes = bytes_per_cand * (W - args.warmup - args.steps / 2.0) + 4.0 * args.k * args.k
    res = {
        "ms": ms, "scored": scored, "value": scored / (ms * 1e-3),
        "us_per_iteration": ms * 1e3 / args.steps,
        "algorithmic_bytes_per_launch": per_iter_bytes,
        "achieved_gbs": per_iter_bytes / (ms * 1e-3 / args.steps) / 1e9,
        "launches": max(m.launches_per_iteration() * args.steps, 2),     # persistent: one launch per select()
        "loop": name, "bytes_per_candidate": bytes_per_cand,
    }
    res["parity_n"] = None if args.skip_parity else mi_parity(args, dist, rank, world, t_pos, t_gain)
    # the same job through the cell-index loop (ACAV_MI_LOOP_CELLS): identical picks, O(K^2) work per iteration
    res["cell_index_loop"] = None
    if not args.skip_cells:
        ms_c, (c_pos, c_gain), mc = time_loop(args, dist, rank, world, cells, "cells", W)
        # both engines have now made warmup + steps picks from the same list: compare picks and tables
        Ns, _, _, sums_s = m.read_state()
        Nc, _, _, sums_c = mc.read_state()
        res["cell_index_loop"] = {
            "us_per_iteration": ms_c * 1e3 / args.steps, "value": scored / (ms_c * 1e-3), "unit": UNIT,
            "loop": mc.loop_name(),
            "same_picks_as_streaming_loop": bool(torch.equal(c_pos, t_pos) and torch.equal(c_gain, t_gain)),
            "same_table_as_streaming_loop": bool(torch.equal(Ns, Nc) and torch.equal(sums_s, sums_c)),
            "what": "same candidate list and iterations through acav_mi_run(ACAV_MI_LOOP_CELLS): candidates sorted once by "
                    "table cell, each iteration scans the K_a x K_v cells instead of the candidates (identical picks)"}
        res["launches"] += 2
        del mc
    # the same loop later in the run: the timed iterations above are picks warmup .. warmup+steps of an EMPTY table, where
    # thousands of cells tie for the best gain; a selection of 1e7 clips spends its life past that
    res["later_iterations"] = None
    if args.later_picks > 0:
        ms_l, _ = timed(dist, lambda: m.select(args.later_picks), lambda: m.select(args.steps))
        m.check_status()
        res["later_iterations"] = {"after_picks": args.warmup + args.steps + args.later_picks,
                                   "us_per_iteration": ms_l * 1e3 / args.steps,
                                   "what": "the same engine, %d more picks on (untimed), then %d timed iterations"
                                           % (args.later_picks, args.steps)}
        res["launches"] += 2
    # the other candidate-stream loop beside it (2 bytes per candidate: round 1's headline kernel)
    res["two_byte_stream_loop"] = None
    if not args.skip_cells and not name.startswith("persistent"):
        ms_p, (p_pos, p_gain), mp = time_loop(args, dist, rank, world, cells, "persistent", W)
        b2 = 2.0 * (W - args.warmup - args.steps / 2.0) + 4.0 * args.k * args.k
        res["two_byte_stream_loop"] = {
            "us_per_iteration": ms_p * 1e3 / args.steps, "value": scored / (ms_p * 1e-3), "unit": UNIT, "loop": mp.loop_name(),
            "achieved_gbs": b2 / (ms_p * 1e-3 / args.steps) / 1e9, "bytes_per_candidate": 2.0,
            "same_picks_as_headline_loop": bool(torch.equal(p_pos, t_pos) and torch.equal(p_gain, t_gain))}
        res["launches"] += 2
        del mp
    del m
    # BASELINE config 4 as written: 1e8 candidates in TOTAL, sharded over the ranks (strong scaling)
    res["c4"] = None
    if not args.skip_strong:
        w_rank = W // world
        if world == 1:
            c4 = {"stream_us_per_iteration": res["us_per_iteration"],
                  "cells_us_per_iteration": (res["cell_index_loop"] or {}).get("us_per_iteration")}
        else:
            ms_s, _, eng = time_loop(args, dist, rank, world, cells, args.mi_loop, w_rank)
            del eng
            ms_cc, _, eng = time_loop(args, dist, rank, world, cells, "cells", w_rank)
            del eng
            c4 = {"stream_us_per_iteration": ms_s * 1e3 / args.steps, "cells_us_per_iteration": ms_cc * 1e3 / args.steps}
            res["launches"] += 4
        best = min(v for v in c4.values() if v is not None)
        c4.update({"candidates_total": w_rank * world, "candidates_per_gpu": w_rank, "scaling": "strong",
                   "projected_seconds_for_1e7_picks": best * 1e-6 * 1e7,
                   "what": "BASELINE config 4 (1e8 candidates in total, sharded over the ranks): microseconds per exact "
                           "greedy iteration for the candidate-stream and the cell-index loop; --workload c4 runs the "
                           "1e7 picks to completion"})
        res["c4"] = c4
    e2e = None
    if not args.skip_e2e:
        host = torch.empty((W, 2), dtype=torch.int64, pin_memory=True)
        host.copy_(cells)
        del cells
        def job():
            torch.cuda.synchronize()
            if dist:
                dist.barrier()
            t0 = time.perf_counter()
            m2 = mi_engine(host, args.k, rank, world, args.mi_loop, W * world, W * rank)
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            from acav100m_b200 import _lib
            try:                                               # the layout build on its own clock (select() would do it)
                _lib.call("acav_mi_prepare", m2._engine, m2._loop_mode(), _lib.stream_ptr(m2.device))
            except _lib.AcavError:
                pass                                           # loop="auto": select() moves on to the next loop itself
            torch.cuda.synchronize()
            t1b = time.perf_counter()
            pos, gain = m2.select(args.steps)
            pos_h, gain_h = pos.cpu(), gain.cpu()
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            del m2
            return t0, t1, t1b, t2, pos_h, gain_h

        # one whole job untimed first (device buffers of the sizes this list needs come out of the library's cache
        # afterwards, as for any user selecting from a second list), then the timed one
        job()
        gc.collect()
        t0, t1, t1b, t2, pos_h, gain_h = job()
        dt = torch.tensor([t2 - t0], device="cuda", dtype=torch.float64)
        if dist:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        scored2 = sum(w_global - i for i in range(args.steps))
        e2e = {"value": scored2 / float(dt.item()), "unit": UNIT,
               "h2d_bytes_per_step": host.numel() * 8 / args.steps,
               "d2h_bytes_per_step": (pos_h.numel() * 8 + gain_h.numel() * 4) / args.steps,
               "seconds": float(dt.item()),
               "phases_ms_rank0": {"h2d_pack_tables": round(1e3 * (t1 - t0), 2),
                                   "layout_build": round(1e3 * (t1b - t1), 2),
                                   "iterations_plus_d2h": round(1e3 * (t2 - t1b), 2)},
               "warmup_jobs": 1,
               "what": "EfficientMemMI built from a pinned host int64 [W,2] tensor + %d greedy iterations + "
                       "D2H of (S, GAIN); one identical job untimed before it" % args.steps}
        res["launches"] += 4
    return res, e2e


# -------------------------------------------------------------------------------------------------
# k-means
# -------------------------------------------------------------------------------------------------

def km_snapshot(km, xb):
    """What the trained model looks like on one batch: clusters in use, skew, rows the bf16 screen could not decide."""
    from acav100m_b200 import _lib
    b, d = xb.shape
    dev = xb.device
    ws = km._workspace(b)
    best = torch.empty(b, dtype=torch.int64, device=dev)
    nref = torch.zeros(2, dtype=torch.int32, device=dev)
    _lib.call("acav_kmeans_assign", ws, _lib.ptr(xb), b, d, _lib.ptr(km.centers), _lib.ptr(km.counts),
              km.underused_threshold(), float(km.reinit[1]), _lib.ptr(best), None, None, _lib.ptr(nref), km._mode(),
              _lib.stream_ptr(dev))
    hist = torch.bincount(best, minlength=km.centers.shape[0])
    return {"clusters_in_use_in_batch": int((hist > 0).sum()), "max_rows_per_centroid": int(hist.max()),
            "rows_rechecked_on_candidates_frac": float(nref[0]) / b, "rows_through_full_exact_kernel_frac": float(nref[1]) / b,
            "underused_centroids": int((km.counts < km.underused_threshold()).sum()), "lr_fallbacks": km.fallback,
            "samples_seen": km.count}


def run_kmeans(args, dist, rank, world, epochs=None, report_pass=True):
    from acav100m_b200 import _lib, synth
    from acav100m_b200.clustering import KMeans
    import types
    dev = torch.device("cuda", torch.cuda.current_device())
    n, d, k, b = args.km_rows, args.km_d, args.k, args.km_batch
    x = synth.gaussian_mixture_torch(n, d, k, 2003 + rank, dev, means_seed=1003)     # shared mixture, per-rank rows
    kargs = types.SimpleNamespace(computation=types.SimpleNamespace(device="cuda", num_gpus=world))
    torch.manual_seed(1003)
    km = KMeans(kargs, d, k, assign_mode=args.km_mode, warmup_rng="cuda",
                graph=False if args.km_graph == "off" else "auto", comm=args.km_comm)
    km.to(dev)
    km.initialize()
    nb = n // b
    steps = args.km_steps or min(args.steps, nb)
    epochs = args.km_epochs if epochs is None else epochs

    def run_steps(cnt, off=0):
        for i in range(cnt):
            j = (off + i) % nb
            km.add(x[j * b:(j + 1) * b], sync=False, distance=False)

    # the state the operator reaches: trained by KMeans.add itself from the reference's init (torch.rand * 1e-5,
    # random-assignment warm-up) over whole epochs of the resident shard with the reference's lr schedule
    # (run_clustering.py:164-176); recurring batch addresses are captured as CUDA graphs on their second visit
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for epoch in range(epochs):
        km.lr = 0.1 ** (2 + epoch // 5)
        run_steps(nb)
    torch.cuda.synchronize()
    train_s = time.perf_counter() - t0
    km.check_status()
    state = km_snapshot(km, x[:b])
    state["epochs_trained"] = epochs
    state["train_seconds"] = train_s
    state["train_iter_per_sec"] = epochs * nb / train_s if train_s > 0 else None

    ms_step, _ = timed(dist, lambda: run_steps(max(args.warmup, 3), 0), lambda: run_steps(steps, max(args.warmup, 3)))
    km.check_status()

    def run_assign(cnt, off=0):
        for i in range(cnt):
            j = (off + i) % nb
            km.calc_best(x[j * b:(j + 1) * b], sync=False, distance=False)

    ms_assign, _ = timed(dist, lambda: run_assign(3), lambda: run_assign(steps, 3))
    pk = peaks()
    out = {
        "metric": "kmeans_iter_per_sec", "value": steps / (ms_step * 1e-3), "unit": "iter/s",
        "global_batch": b * world, "k": k, "d": d, "rows_resident_per_gpu": n, "steps": steps,
        "ms_per_step": ms_step / steps, "samples_per_sec": steps * b * world / (ms_step * 1e-3),
        "assign_rows_per_sec": steps * b * world / (ms_assign * 1e-3),
        "assign_ms_per_batch": ms_assign / steps, "assign_mode": km.mode_name(),
        "batch_assign_tflops": 2.0 * b * k * d / (ms_assign * 1e-3 / steps) / 1e12,
        "graph_replay": bool(km._gs and km._gs["graphs"]), "exchange": km.comm_name() if world > 1 else None,
        "state": state,
        "state_note": "model trained by KMeans.add from the reference init for %d epochs of the resident shard "
                      "(lr schedule of run_clustering.py:168) before timing" % epochs,
        "gpu_launches": km.launches_per_step() * steps,
    }
    if report_pass:
        # whole assignment pass over the resident shard (KMeans.assign_all: 131072-row chunks, the fp32->bf16
        # preparation of chunk i+1 overlapped with the tensor-core kernel of chunk i)
        ms_pass, _ = timed(dist, lambda: km.assign_all(x[:262144]), lambda: km.assign_all(x))
        # the dominant kernel by itself: the tcgen05 distance GEMM (+ classification) over the whole shard on operands
        # prepared beforehand, chunk by chunk through the C ABI (acav_kmeans_assign_prepared)
        chunk = 131072
        ws = km._workspace(chunk)
        thr, rr = km.underused_threshold(), float(km.reinit[1])
        st = _lib.stream_ptr(dev)
        best_all = torch.empty(n, dtype=torch.int64, device=dev)
        _lib.call("acav_kmeans_prepare_centers", ws, _lib.ptr(km.centers), _lib.ptr(km.counts), thr, rr, st)
        ms_gemm, n_gemm_launch = 0.0, 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for rep in range(2):                                   # first repetition warms up
            ms_gemm, n_gemm_launch = 0.0, 0
            for lo in range(0, n, chunk):
                xb = x[lo:lo + chunk]
                _lib.call("acav_kmeans_prepare_batch", ws, _lib.ptr(xb), xb.shape[0], d, st)
                e0.record()
                _lib.call("acav_kmeans_assign_prepared", ws, _lib.ptr(xb), xb.shape[0], d, _lib.ptr(km.centers),
                          _lib.ptr(km.counts), thr, rr, _lib.c_vp(best_all.data_ptr() + 8 * lo), None, None, None, st)
                e1.record()
                e1.synchronize()
                ms_gemm += e0.elapsed_time(e1)
                n_gemm_launch += 1
        t_ms = torch.tensor([ms_gemm], device="cuda", dtype=torch.float64)
        if dist:
            dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        ms_gemm = float(t_ms.item())
        pass_tflops = 2.0 * n * k * d / (ms_pass * 1e-3) / 1e12
        gemm_tflops = 2.0 * n * k * d / (ms_gemm * 1e-3) / 1e12
        traffic = ncu_traffic("km_assign_pair_kernel")
        out.update({
            "assign_pass_ms": ms_pass, "assign_pass_rows_per_sec": n * world / (ms_pass * 1e-3),
            "assign_pass_tflops": pass_tflops, "assign_pass_frac_of_burst_peak": pass_tflops / pk["bf16_tflops"],
            "roofline": {"bound": "tensor", "achieved": gemm_tflops, "peak": pk["bf16_tflops"],
                         "unit": "TFLOP/s", "frac": gemm_tflops / pk["bf16_tflops"],
                         "frac_of_sustained_peak": gemm_tflops / pk["bf16_tflops_sustained"] if pk.get("bf16_tflops_sustained") else None,
                         "traffic": traffic.get("dram_bytes_per_launch") if traffic else None,
                         "traffic_source": traffic.get("source") if traffic else None,
                         "kernel": "km_assign_pair_kernel<1> (tcgen05 cta_group::2 distance GEMM, 256x256 pair tiles) + top-4 "
                                   "classification + exact re-check, per 131072-row launch; 2*b*K*D flop; trained state",
                         "launches_timed": n_gemm_launch, "ms_per_launch": ms_gemm / max(n_gemm_launch, 1),
                         "algorithmic_flops_per_launch": 2.0 * chunk * k * d,
                         "peak_source": pk["source"] + " bf16 burst (cuBLAS 8192^3)"}})
    if not args.skip_e2e:
        host = torch.empty((steps, b, d), dtype=torch.float32).pin_memory()
        host.copy_(x[:steps * b].view(steps, b, d))
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        copy_stream = torch.cuda.Stream(device=dev)
        cur = torch.cuda.current_stream(dev)
        bufs = [torch.empty((b, d), dtype=torch.float32, device=dev) for _ in range(2)]

        def e2e_pass():
            ready = [torch.cuda.Event() for _ in range(2)]
            freed = [torch.cuda.Event() for _ in range(2)]
            with torch.cuda.stream(copy_stream):
                bufs[0].copy_(host[0], non_blocking=True)
                ready[0].record(copy_stream)
            last = None
            for i in range(steps):
                if i + 1 < steps:                              # H2D of batch i+1 overlaps the step on batch i
                    j = (i + 1) % 2
                    with torch.cuda.stream(copy_stream):
                        if i >= 1:
                            copy_stream.wait_event(freed[j])
                        bufs[j].copy_(host[i + 1], non_blocking=True)
                        ready[j].record(copy_stream)
                cur.wait_event(ready[i % 2])
                last = km.add(bufs[i % 2], sync=False)
                freed[i % 2].record(cur)
            return float(last)

        e2e_pass()                                             # the two staging buffers' graphs are captured here
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        t0 = time.perf_counter()
        e2e_pass()
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
        if dist:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        out["e2e"] = {"value": steps / float(dt.item()), "unit": "iter/s", "h2d_bytes_per_step": b * d * 4,
                      "d2h_bytes_per_step": 4, "samples_per_sec": steps * b * world / float(dt.item()),
                      "what": "KMeans.add on pinned host batches (H2D of batch i+1 on a copy stream while step i runs), "
                              "mean distance read back"}
    del x
    return out


# -------------------------------------------------------------------------------------------------
# BASELINE configs 2, 3, 4 end to end (selectable workloads)
# -------------------------------------------------------------------------------------------------

def run_config4(args, dist, rank, world):
    """C4 as written: 1e8 precomputed cluster-id pairs in total -> 1e7 picks, exact, to completion."""
    from acav100m_b200 import synth
    W_total, picks = args.mi_candidates, args.c4_picks
    w_rank = W_total // world
    dev = torch.device("cuda", torch.cuda.current_device())
    cells = synth.zipf_pairs_torch(w_rank, args.k, 1004 + rank, dev)     # contiguous rank ranges of one global list
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    t0 = time.perf_counter()
    m = mi_engine(cells, args.k, rank, world, "auto", w_rank * world, w_rank * rank, max_picks=picks + 16)   # auto -> cell index at K = 1024
    chunk, done, pos_all, gain_all = 1_000_000, 0, [], []
    marks = []
    while done < picks:
        n = min(chunk, picks - done)
        p, g = m.select(n)
        pos_all.append(p)
        gain_all.append(g)
        done += n
        torch.cuda.synchronize()
        marks.append((done, round(time.perf_counter() - t0, 3)))
    m.check_status()
    wall = time.perf_counter() - t0
    pos = torch.cat(pos_all)
    gain = torch.cat(gain_all)
    # checks: (1) picks are distinct; (2) the first picks equal the C oracle's on the whole list (rank 0 regenerates every
    # rank's range); (3) the final table is the histogram of the picked cells
    N, _, _, sums = m.read_state()
    ok_distinct = bool(torch.unique(pos).numel() == pos.numel())
    check_n = min(20_000, picks)
    oracle_ok = table_ok = None
    if rank == 0:
        whole = np.concatenate([synth.zipf_pairs_torch(w_rank, args.k, 1004 + r, dev).cpu().numpy() for r in range(world)])
        want = cpu_oracle_picks(whole, args.k, check_n)
        oracle_ok = bool(np.array_equal(pos[:check_n].cpu().numpy(), want[0]) and
                         np.array_equal(gain[:check_n].cpu().numpy(), want[1]))
        picked = whole[pos.cpu().numpy()]
        want_N = np.zeros((args.k, args.k), dtype=np.int64)
        np.add.at(want_N, (picked[:, 0], picked[:, 1]), 1)
        table_ok = bool(np.array_equal(N.numpy(), want_N))
    return {"config": "C4: %d candidate pairs in total over %d GPU(s), K=%d, %d exact greedy picks" % (w_rank * world, world, args.k, picks),
            "loop": m.loop_name(), "wall_seconds": wall, "us_per_iteration": wall * 1e6 / picks,
            "candidate_clips_per_sec": (w_rank * world - picks / 2.0) * picks / wall,
            "progress_picks_seconds": marks[:: max(len(marks) // 10, 1)], "last_gain": float(gain[-1]),
            "picks_distinct": ok_distinct, "first_%d_picks_equal_c_oracle" % check_n: oracle_ok,
            "final_table_is_histogram_of_picks": table_ok, "n": float(sums[3])}


def run_config3(args, dist, rank, world):
    """C3: 10 M x 2048 rows over 8 GPUs (1.25 M per GPU), K = 1024: one training epoch + the assignment pass."""
    km = run_kmeans(args, dist, rank, world, epochs=1, report_pass=True)
    nb = args.km_rows // args.km_batch
    return {"config": "C3: %d x %d rows over %d GPU(s), K=%d, global batch %d: one epoch (%d steps) + assignment pass"
                      % (args.km_rows * world, args.km_d, world, args.k, args.km_batch * world, nb),
            "epoch_seconds": km["state"]["train_seconds"], "epoch_iter_per_sec": km["state"]["train_iter_per_sec"],
            "assign_pass_ms": km["assign_pass_ms"], "assign_pass_tflops_per_gpu": km["assign_pass_tflops"],
            "assign_pass_frac_of_burst_peak": km["assign_pass_frac_of_burst_peak"],
            "steady_state_iter_per_sec": km["value"], "state": km["state"], "exchange": km["exchange"],
            "gemm_roofline": km["roofline"]}


def run_config2(args, dist, rank, world):
    """C2 on one GPU: 1 M clips, D_a = 512, D_v = 2048, K = 256: k-means on both modalities (one epoch at batch
    65 536 + assignment pass), then exact greedy selection of 100 K clips from the resulting id pairs."""
    from acav100m_b200 import synth
    from acav100m_b200.clustering import KMeans
    dev = torch.device("cuda", torch.cuda.current_device())
    n, k, b = 1_000_000, 256, 65_536
    out = {"config": "C2: 1e6 clips, D_a=512, D_v=2048, K=256, select 1e5, one B200"}
    ids = []
    t_all = time.perf_counter()
    for name, d, seed in (("audio", 512, 2001), ("visual", 2048, 2002)):
        x = synth.gaussian_mixture_torch(n, d, k, seed, dev)
        torch.manual_seed(seed)
        km = KMeans(None, d, k, warmup_rng="cuda")
        km.to(dev)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        km.lr = 1e-2
        for j in range(n // b):
            km.add(x[j * b:(j + 1) * b], sync=False, distance=False)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        best = km.assign_all(x)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        ids.append(best)
        out["kmeans_" + name] = {"d": d, "train_s": t1 - t0, "steps": n // b, "assign_s": t2 - t1,
                                 "assign_tflops": 2.0 * n * k * d / (t2 - t1) / 1e12,
                                 "assign_gbs_fp32_rows": n * d * 4 / (t2 - t1) / 1e9,
                                 "clusters_used": int(torch.unique(best).numel()), "lr_fallbacks": km.fallback}
        del x, km
    cells = torch.stack(ids, dim=1).contiguous()
    t0 = time.perf_counter()
    m = mi_engine(cells, k, 0, 1, "auto", n, 0)
    picks = 100_000 - 2
    pos, gain = m.select(picks)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    out["mi"] = {"loop": m.loop_name(), "select_s": t1 - t0, "picks": picks, "us_per_iteration": (t1 - t0) * 1e6 / picks,
                 "last_gain": float(gain[-1]), "first_picks": pos[:5].tolist()}
    out["total_s"] = time.perf_counter() - t_all
    return out


# -------------------------------------------------------------------------------------------------
# CPU arms (oracle port of the reference; the only place bench.py touches oracle/)
# -------------------------------------------------------------------------------------------------

def cpu_mi(args, repeats, warm=1):
    from acav100m_b200 import synth
    from oracle import mi_oracle as mo
    W = min(args.cpu_sample, args.mi_candidates)
    a = synth.zipf_pairs(W, args.k, 1004)
    threads = os.cpu_count() or 1
    c1, c2 = a[:, 0].astype(np.int32), a[:, 1].astype(np.int32)
    mo.scan_once_seconds(c1, c2, args.k, max(warm, 1), threads)      # warm-up scans
    sec = mo.scan_once_seconds(c1, c2, args.k, repeats, threads)
    return {"value": W * repeats / sec, "unit": UNIT, "cores": threads, "kind": "port", "seconds": sec,
            "sample": "%d full scans of a %d-candidate sample (same Zipf pair distribution, K=%d) with "
                      "oracle/mi_oracle.c, OpenMP over %d threads" % (repeats, W, args.k, threads)}


def cpu_mi_torch(args, seconds=8.0):
    """The reference's own formulation -- EfficientMemMI's torch ops ([P, W, C] gathers per iteration, mi.py:322-381)
    as restated in oracle/mi_oracle.py::greedy_mem_mi -- on a 1e6-candidate prefix, all host threads."""
    from acav100m_b200 import synth
    from oracle import mi_oracle as mo
    torch.set_num_threads(os.cpu_count() or 1)
    W, k = 1_000_000, min(args.k, 256)                        # [W, C] fp32 gathers: 1 GB per iteration at C = 256
    a = synth.zipf_pairs(W, k, 1004)
    iters, t0 = 0, time.perf_counter()
    subset = 4
    while True:
        mo.greedy_mem_mi(a, k, [(0, 1)], list(range(1, W)), subset, [0])
        iters += subset - 2
        if time.perf_counter() - t0 > seconds:
            break
        subset += 2
    sec = time.perf_counter() - t0
    return {"value": W * iters / sec, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port-torch", "seconds": sec,
            "sample": "%d greedy iterations over a %d-candidate prefix at K=%d with the reference's torch formulation "
                      "(oracle/mi_oracle.py::greedy_mem_mi, incl. its per-run setup)" % (iters, W, k)}


def cpu_oracle_picks(cells, k, picks):
    """The C oracle's greedy picks and fp32 gains on a whole candidate list (checker of mi_parity / config 4)."""
    from oracle import mi_oracle as mo
    return mo.greedy_mem_mi_c(cells[:, 0], cells[:, 1], k, picks, bucketed=True)


def cpu_kmeans(args, steps):
    from acav100m_b200 import synth
    from oracle import kmeans_oracle as ko
    torch.set_num_threads(os.cpu_count() or 1)
    d, k, b = args.km_d, args.k, args.km_batch
    x = torch.from_numpy(synth.gaussian_mixture(b * 2, d, k, 1003))
    st = ko.new_state(d, k)
    st.centers = x[:k].clone() if k <= len(x) else torch.from_numpy(synth.gaussian_mixture(k, d, k, 1))
    st.count = 10 * k * 8
    ko.sgd_step(st, x[:b])
    t0 = time.perf_counter()
    for i in range(steps):
        ko.sgd_step(st, x[(i % 2) * b:(i % 2 + 1) * b])
    sec = time.perf_counter() - t0
    return {"value": steps / sec, "unit": "iter/s", "cores": torch.get_num_threads(), "kind": "port",
            "samples_per_sec": steps * b / sec,
            "sample": "%d KMeans.add steps, b=%d d=%d k=%d, torch CPU (oracle/kmeans_oracle.py)" % (steps, b, d, k)}


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def config_dict(args, world):
    return {"workload": "C4 candidate list (100M Zipf cluster-id pairs, K=1024) per GPU for greedy MI; "
                        "C3 per-GPU shard (1.25M x 2048 fp32, K=1024, batch 8192/GPU) for k-means",
            "mi_candidates_per_gpu": args.mi_candidates, "k": args.k, "km_rows_per_gpu": args.km_rows,
            "km_d": args.km_d, "km_batch_per_gpu": args.km_batch, "parallelism": "shard%d" % world,
            "l2": ("greedy MI: every iteration streams the whole candidate list of the GPU -- %.0f MB in the one-byte layout "
                   "(--mi-loop bytes, loaded with an L2 evict-first policy: ncu counts 104 MB of DRAM reads per iteration at "
                   "1e8 candidates, L2 hit rate 13 %%, profiles/r02_mi_bytes_final.ncu.txt and roofline.traffic -- nothing is "
                   "served from the 126 MB L2), %.0f MB > L2 in the 2-byte layout (two_byte_stream_loop); k-means walks "
                   "%.1f GB/GPU of resident rows"
                   % (args.mi_candidates / 1e6, args.mi_candidates * 2 / 1e6, args.km_rows * args.km_d * 4 / 1e9))}


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.perf_counter()
    mi = cpu_mi(args, args.steps, args.warmup)        # one step = one full scan of the bounded sample
    mi["torch_formulation"] = cpu_mi_torch(args, 6.0)
    km = None if args.skip_kmeans else cpu_kmeans(args, max(min(args.steps, 20), 3))
    W = min(args.cpu_sample, args.mi_candidates)
    line = {
        "impl": "reference", "metric": METRIC, "value": mi["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * W / mi["value"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(args, args.gpus), "cpu_baseline": mi, "cpu_model": cpu_model(),
        "e2e": {"value": mi["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "kmeans": km, "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    if args.impl == "reference":
        return main_reference(args)
    from acav100m_b200 import _lib
    _lib.load()
    dist, rank, world, local = dist_setup(args.gpus)
    sampler = ClockSampler(local)
    sampler.start()
    if args.workload != "default":
        fn = {"c2": run_config2, "c3": run_config3, "c4": run_config4}[args.workload]
        res = fn(args, dist, rank, world)
        res["clocks"] = sampler.stop()
        res["n_gpus"] = world
        if rank == 0:
            print(json.dumps({"workload": args.workload, **res}), flush=True)
        if dist:
            dist.barrier()
            dist.destroy_process_group()
        return
    if args.skip_mi:
        km = run_kmeans(args, dist, rank, world)
        if rank == 0:
            print(json.dumps({"kmeans": km}), flush=True)
        if dist:
            dist.barrier()
            dist.destroy_process_group()
        return
    mi, e2e = run_mi(args, dist, rank, world)
    km = None if args.skip_kmeans else run_kmeans(args, dist, rank, world)
    clocks = sampler.stop()
    if rank == 0:
        pk = peaks()
        cpu = None
        if world == 1 and not args.skip_cpu_baseline:
            cpu = cpu_mi(args, 2000, 3)                    # ~10 s of host work on 16 threads (bounded sample)
            cpu["torch_formulation"] = cpu_mi_torch(args, 8.0)
            if km is not None:
                km["cpu_baseline"] = cpu_kmeans(args, 200)  # ~5 s
        traffic = ncu_traffic("mi_persistent_kernel" if mi["loop"].startswith("persistent") else "mi_stream8_kernel")
        roofline = {"bound": "hbm", "achieved": mi["achieved_gbs"], "peak": pk["hbm_gbs"], "unit": "GB/s",
                    "frac": mi["achieved_gbs"] / pk["hbm_gbs"],
                    "traffic": traffic.get("dram_bytes_per_launch") if traffic else None,
                    "traffic_source": traffic.get("source") if traffic else None,
                    "bytes_per_candidate_accounted": mi["bytes_per_candidate"],
                    "list_order_equivalent_gbs": mi["achieved_gbs"] * 4.0 / mi["bytes_per_candidate"],
                    "kernel": "greedy-MI iteration (gain rows + candidate stream scan + winner hand-over), per GPU; "
                              "one launch = one iteration of the persistent kernel",
                    "algorithmic_bytes_per_launch": mi["algorithmic_bytes_per_launch"],
                    "peak_source": pk["source"] + " copy bandwidth"}
        if km is not None:
            # the driver's record keeps `roofline` and `e2e` whole: the k-means numbers ride along here
            roofline["kmeans_step_iter_per_sec"] = km["value"]
            roofline["kmeans_ms_per_step"] = km["ms_per_step"]
            roofline["kmeans_samples_per_sec"] = km["samples_per_sec"]
            roofline["kmeans_assign_pass_ms"] = km.get("assign_pass_ms")
            roofline["kmeans_assign_pass_frac_of_burst_peak"] = km.get("assign_pass_frac_of_burst_peak")
            roofline["kmeans_assign_gemm"] = km.get("roofline")
            roofline["kmeans_state"] = km["state"]
            if e2e is not None and km.get("e2e"):
                e2e["kmeans"] = km["e2e"]
        if mi.get("c4"):
            roofline["c4_strong_scaling"] = mi["c4"]
        if mi.get("cell_index_loop"):
            roofline["cell_index_loop_us_per_iteration"] = mi["cell_index_loop"]["us_per_iteration"]
        if mi.get("two_byte_stream_loop"):
            t2 = mi["two_byte_stream_loop"]
            roofline["two_byte_stream_loop"] = {"us_per_iteration": t2["us_per_iteration"], "achieved": t2["achieved_gbs"],
                                                "frac": t2["achieved_gbs"] / pk["hbm_gbs"], "unit": "GB/s",
                                                "bytes_per_candidate_accounted": 2.0}
        line = {
            "metric": METRIC, "value": mi["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": mi["ms"] / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(args, world), "us_per_iteration": mi["us_per_iteration"],
            "mi_loop": mi["loop"], "gpu_launches": mi["launches"] + (km["gpu_launches"] if km else 0),
            "e2e": e2e, "roofline": roofline,
            "parity_n": (mi["parity_n"] or {}).get("result"), "parity": mi["parity_n"],
            "later_iterations": mi.get("later_iterations"),
            "cell_index_loop": mi.get("cell_index_loop"), "two_byte_stream_loop": mi.get("two_byte_stream_loop"),
            "c4": mi.get("c4"),
            "cpu_baseline": cpu, "cpu_model": cpu_model(), "kmeans": km, "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if dist:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
