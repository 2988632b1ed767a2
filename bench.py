#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on B200: greedy-MI candidate-clips/s (+ k-means iter/s), K = 1024.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config.workload, DESIGN.md "Measurement"):
  * greedy MI: BASELINE config 4's candidate list -- 100 M (c_a, c_v) cluster-id pairs, K_a = K_v =
    1024 -- PER GPU (weak scaling; larger than L2, so every iteration streams from HBM).  One step =
    one exact greedy iteration: score every remaining candidate, first arg-max, table update, removal.
    value = candidates scored per second over the K timed steps, summed over ranks.
  * k-means (reported under "kmeans"): BASELINE config 3's per-GPU shard -- 1.25 M x 2048 fp32 rows
    resident in HBM, K = 1024, per-GPU batch 8192 (global 65 536 at 8 GPUs) -- SGD steps/s via
    KMeans.add and assignment rows/s via KMeans.calc_best.
Inputs are synthetic (acav100m_b200.synth) and resident in HBM when the timed region starts; "e2e"
repeats the measurement through the public API starting from pinned HOST buffers.
`--impl reference` times the reference's CPU implementation (the oracle port; the reference is pure
Python and does not travel to the GPU box) on the host cores for the same metric.
"""
import argparse
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
    p.add_argument("--k", type=int, default=1024)
    p.add_argument("--mi-candidates", type=int, default=100_000_000, help="per GPU")
    p.add_argument("--mi-loop", default="auto")
    p.add_argument("--km-rows", type=int, default=1_250_000, help="per GPU, resident")
    p.add_argument("--km-d", type=int, default=2048)
    p.add_argument("--km-batch", type=int, default=8192, help="per GPU")
    p.add_argument("--km-steps", type=int, default=0, help="0 = min(steps, rows/batch)")
    p.add_argument("--km-mode", default="auto")
    p.add_argument("--cpu-sample", type=int, default=10_000_000)
    p.add_argument("--skip-cpu-baseline", action="store_true")
    p.add_argument("--skip-kmeans", action="store_true")
    p.add_argument("--skip-mi", action="store_true", help="development only: the line then has no headline value")
    p.add_argument("--skip-e2e", action="store_true")
    p.add_argument("--skip-cells", action="store_true")
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
# our arm
# -------------------------------------------------------------------------------------------------

def mi_engine(cells_src, k, rank, world, loop, w_global, lo):
    from acav100m_b200.subset_selection import get_measure
    shard = (rank, world) if world > 1 else None
    m = get_measure("mem_mi")(cells_src, ncentroids=k, device="cuda", shard=shard, loop=loop)
    m.init_from_cells([(0, 1)], cells_src, w_global=w_global, lo=lo)
    return m


def mi_parity(args, dist, rank, world, timed_pos, timed_gain):
    """Self-certification of a (multi-GPU) run, printed as `parity_n` in the JSON line:
      1. every rank returned the same (positions, gains) from the timed iterations;
      2. a 2e6-candidates-per-rank list, sharded exactly like the timed one, run through the same loop(s), gives the
         picks and fp32 gains of the C oracle (oracle/mi_oracle.c, bucketed scan) on the WHOLE list, bit for bit.
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
    for loop in ("persistent", "cells"):
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


def run_mi(args, dist, rank, world):
    from acav100m_b200 import synth
    W = args.mi_candidates
    dev = torch.device("cuda", torch.cuda.current_device())
    cells = synth.zipf_pairs_torch(W, args.k, 1004 + rank, dev)
    m = mi_engine(cells, args.k, rank, world, args.mi_loop, W * world, W * rank)
    ms, (t_pos, t_gain) = timed(dist, lambda: m.select(args.warmup), lambda: m.select(args.steps))
    m.check_status()
    w_global = W * world
    scored = sum(w_global - args.warmup - i for i in range(args.steps))
    # algorithmic bytes per iteration and GPU in the layout the loop actually streams:
    # persistent = 2-byte c2 stream + 4*K^2 table counts; list order = 4-byte packed pairs + 4*K^2 gains
    bytes_per_cand = 2.0 if m.loop_name().startswith("persistent") else 4.0
    per_iter_bytes = bytes_per_cand * (W - args.warmup - args.steps / 2.0) + 4.0 * args.k * args.k
    res = {
        "ms": ms, "scored": scored, "value": scored / (ms * 1e-3),
        "us_per_iteration": ms * 1e3 / args.steps,
        "algorithmic_bytes_per_launch": per_iter_bytes,
        "achieved_gbs": per_iter_bytes / (ms * 1e-3 / args.steps) / 1e9,
        "launches": max(m.launches_per_iteration() * args.steps, 2),     # persistent: one launch per select()
        "loop": m.loop_name(), "bytes_per_candidate": bytes_per_cand,
    }
    res["parity_n"] = None if args.skip_parity else mi_parity(args, dist, rank, world, t_pos, t_gain)
    # the same job through the cell-index loop (ACAV_MI_LOOP_CELLS): identical picks, O(K^2) work per iteration
    res["cell_index_loop"] = None
    if not args.skip_cells and world == 1:
        mc = mi_engine(cells, args.k, rank, world, "cells", W * world, W * rank)
        ms_c, _ = timed(dist, lambda: mc.select(args.warmup), lambda: mc.select(args.steps))
        # both engines have now made warmup + steps picks from the same list: compare their tables
        Ns, _, _, sums_s = m.read_state()
        Nc, _, _, sums_c = mc.read_state()
        res["cell_index_loop"] = {
            "us_per_iteration": ms_c * 1e3 / args.steps, "value": scored / (ms_c * 1e-3), "unit": UNIT,
            "same_table_as_streaming_loop": bool(torch.equal(Ns, Nc) and torch.equal(sums_s, sums_c)),
            "what": "same candidate list and iterations through acav_mi_run(ACAV_MI_LOOP_CELLS): candidates sorted once by "
                    "table cell, each iteration scans the K_a x K_v cells instead of the candidates (identical picks)"}
        del mc
    e2e = None
    if not args.skip_e2e:
        del m
        host = torch.empty((W, 2), dtype=torch.int64, pin_memory=True)
        host.copy_(cells)
        del cells
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        t0 = time.perf_counter()
        m2 = mi_engine(host, args.k, rank, world, args.mi_loop, W * world, W * rank)
        pos, gain = m2.select(args.steps)
        pos_h, gain_h = pos.cpu(), gain.cpu()
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
        if dist:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        scored2 = sum(w_global - i for i in range(args.steps))
        e2e = {"value": scored2 / float(dt.item()), "unit": UNIT,
               "h2d_bytes_per_step": host.numel() * 8 / args.steps,
               "d2h_bytes_per_step": (pos_h.numel() * 8 + gain_h.numel() * 4) / args.steps,
               "seconds": float(dt.item()),
               "what": "EfficientMemMI built from a pinned host int64 [W,2] tensor + %d greedy iterations + "
                       "D2H of (S, GAIN)" % args.steps}
    return res, e2e


def run_kmeans(args, dist, rank, world):
    from acav100m_b200 import synth
    from acav100m_b200.clustering import KMeans
    import types
    dev = torch.device("cuda", torch.cuda.current_device())
    n, d, k, b = args.km_rows, args.km_d, args.k, args.km_batch
    x = synth.gaussian_mixture_torch(n, d, k, 1003 + rank, dev)
    kargs = types.SimpleNamespace(computation=types.SimpleNamespace(device="cuda", num_gpus=world))
    torch.manual_seed(1003)
    km = KMeans(kargs, d, k, assign_mode=args.km_mode, warmup_rng="cuda")
    km.to(dev)
    km.initialize()
    km.lr = 1e-2
    nb = n // b
    steps = args.km_steps or min(args.steps, nb)
    warm = min(args.warmup, nb)

    def run_steps(cnt, off=0):
        for i in range(cnt):
            j = (off + i) % nb
            km.add(x[j * b:(j + 1) * b], sync=False)

    # train from the reference's init through its warm-up (random assignment until 10*k samples) and a
    # few dozen SGD steps, so the timed steps see the operator in its steady state
    settle = -(-10 * k // (b * world)) + 24
    run_steps(settle)
    warm += settle

    ms_step, _ = timed(dist, lambda: run_steps(3, warm - 3), lambda: run_steps(steps, warm))

    def run_assign(cnt, off=0):
        for i in range(cnt):
            j = (off + i) % nb
            km.calc_best(x[j * b:(j + 1) * b], sync=False)

    ms_assign, _ = timed(dist, lambda: run_assign(3), lambda: run_assign(steps, 3))
    # the same operator on a converged model: centroids at the mixture's component means (every row has one
    # clear nearest centroid, so the tensor-core screen decides it without the exact re-check)
    km_sep = KMeans(kargs, d, k, assign_mode=args.km_mode, warmup_rng="cuda")
    km_sep.to(dev)
    gm = torch.Generator(device=dev).manual_seed(1003 + rank)
    means = torch.randn(k, d, generator=gm, device=dev) * 3.0          # the means gaussian_mixture_torch drew
    km_sep.centers.copy_(means)
    km_sep.counts.fill_(1000.0)
    km_sep.count = 1000 * k
    km_sep.lr = 1e-3

    def sep_steps(cnt, off=0):
        for i in range(cnt):
            j = (off + i) % nb
            km_sep.add(x[j * b:(j + 1) * b], sync=False)

    def sep_assign(cnt, off=0):
        for i in range(cnt):
            j = (off + i) % nb
            km_sep.calc_best(x[j * b:(j + 1) * b], sync=False)

    ms_sep_step, _ = timed(dist, lambda: sep_steps(3), lambda: sep_steps(steps, 3))
    ms_sep_assign, _ = timed(dist, lambda: sep_assign(3), lambda: sep_assign(steps, 3))
    # whole assignment pass over the resident shard (KMeans.assign_all: 131072-row chunks, the fp32->bf16
    # preparation of chunk i+1 overlapped with the tensor-core kernel of chunk i)
    ms_pass, _ = timed(dist, lambda: km_sep.assign_all(x[:262144]), lambda: km_sep.assign_all(x))
    # the dominant kernel by itself: the tcgen05 distance GEMM (+ classification) over the whole shard on operands
    # prepared beforehand, chunk by chunk through the C ABI (acav_kmeans_assign_prepared)
    from acav100m_b200 import _lib
    chunk = 131072
    ws = km_sep._workspace(chunk)
    thr, rr = km_sep.underused_threshold(), float(km_sep.reinit[1])
    st = _lib.stream_ptr(dev)
    best_all = torch.empty(n, dtype=torch.int64, device=dev)
    _lib.call("acav_kmeans_prepare_centers", ws, _lib.ptr(km_sep.centers), _lib.ptr(km_sep.counts), thr, rr, st)
    ms_gemm, n_gemm_launch = 0.0, 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for rep in range(2):                                   # first repetition warms up
        ms_gemm, n_gemm_launch = 0.0, 0
        for lo in range(0, n, chunk):
            xb = x[lo:lo + chunk]
            _lib.call("acav_kmeans_prepare_batch", ws, _lib.ptr(xb), xb.shape[0], d, st)
            e0.record()
            _lib.call("acav_kmeans_assign_prepared", ws, _lib.ptr(xb), xb.shape[0], d, _lib.ptr(km_sep.centers),
                      _lib.ptr(km_sep.counts), thr, rr, _lib.c_vp(best_all.data_ptr() + 8 * lo), None, None, None, st)
            e1.record()
            e1.synchronize()
            ms_gemm += e0.elapsed_time(e1)
            n_gemm_launch += 1
    t_ms = torch.tensor([ms_gemm], device="cuda", dtype=torch.float64)
    if dist:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_gemm = float(t_ms.item())
    pk = peaks()
    flops = 2.0 * b * k * d
    t_assign = ms_sep_assign * 1e-3 / steps
    pass_tflops = 2.0 * n * k * d / (ms_pass * 1e-3) / 1e12
    gemm_tflops = 2.0 * n * k * d / (ms_gemm * 1e-3) / 1e12
    out = {
        "metric": "kmeans_iter_per_sec", "value": steps / (ms_step * 1e-3), "unit": "iter/s",
        "global_batch": b * world, "k": k, "d": d, "rows_resident_per_gpu": n, "steps": steps,
        "ms_per_step": ms_step / steps, "samples_per_sec": steps * b * world / (ms_step * 1e-3),
        "assign_rows_per_sec": steps * b * world / (ms_assign * 1e-3),
        "assign_ms_per_batch": ms_assign / steps, "assign_mode": km.mode_name(),
        "steps_before_timing": warm, "lr_fallbacks": km.fallback,
        "state": "trained from the reference init (torch.rand*1e-5, random-assignment warm-up): an early, skewed state in "
                 "which a handful of centroids own the batch and about half of the rows have 2-16 possible winners "
                 "after the bf16 screen (exact candidate re-check)",
        "converged": {
            "state": "centroids at the mixture means (one clear nearest centroid per row)",
            "value": steps / (ms_sep_step * 1e-3), "unit": "iter/s", "ms_per_step": ms_sep_step / steps,
            "samples_per_sec": steps * b * world / (ms_sep_step * 1e-3),
            "assign_rows_per_sec": steps * b * world / (ms_sep_assign * 1e-3),
            "assign_ms_per_batch": ms_sep_assign / steps,
            "batch_assign_tflops": flops / t_assign / 1e12,
            "assign_pass_ms": ms_pass, "assign_pass_rows_per_sec": n * world / (ms_pass * 1e-3),
            "assign_pass_tflops": pass_tflops,
            "assign_pass_note": "fp32->bf16 preparation (HBM-bound, ~980 W) and the GEMM (~930 W) both run at the "
                                "board power limit, so the pass is the sum of the two (DESIGN.md 2.4)"},
        "gpu_launches": km.launches_per_step() * steps,
        "roofline": {"bound": "tensor", "achieved": gemm_tflops, "peak": pk["bf16_tflops"],
                     "unit": "TFLOP/s", "frac": gemm_tflops / pk["bf16_tflops"],
                     "frac_of_sustained_peak": gemm_tflops / pk["bf16_tflops_sustained"] if pk.get("bf16_tflops_sustained") else None,
                     "traffic": 613.3e6 + 11.4e6,
                     "traffic_source": "ncu dram__bytes_read+write per launch (131072 x 2048 bf16 rows = 537 MB + centroids), "
                                       "profiles/r01_km_pair256.ncu.txt",
                     "kernel": "km_assign_pair_kernel<1> (tcgen05 cta_group::2 distance GEMM, 256x256 pair tiles) + top-4 "
                               "classification, per 131072-row launch; 2*b*K*D flop; converged state",
                     "launches_timed": n_gemm_launch, "ms_per_launch": ms_gemm / max(n_gemm_launch, 1),
                     "algorithmic_flops_per_launch": 2.0 * chunk * k * d,
                     "tensor_pipe_active_pct_ncu": 81.5,
                     "peak_source": pk["source"] + " bf16 burst (cuBLAS 8192^3)"},
    }
    if not args.skip_e2e:
        host = torch.empty((steps, b, d), dtype=torch.float32).pin_memory()
        host.copy_(x[:steps * b].view(steps, b, d))
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        t0 = time.perf_counter()
        last = None
        copy_stream = torch.cuda.Stream(device=dev)
        cur = torch.cuda.current_stream(dev)
        bufs = [torch.empty((b, d), dtype=torch.float32, device=dev) for _ in range(2)]
        ready = [torch.cuda.Event() for _ in range(2)]
        freed = [torch.cuda.Event() for _ in range(2)]
        with torch.cuda.stream(copy_stream):
            bufs[0].copy_(host[0], non_blocking=True)
            ready[0].record(copy_stream)
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
        float(last)
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
        if dist:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        out["e2e"] = {"value": steps / float(dt.item()), "unit": "iter/s", "h2d_bytes_per_step": b * d * 4,
                      "d2h_bytes_per_step": 4, "what": "KMeans.add on pinned host batches (H2D of batch i+1 on a copy stream while step i runs), mean distance read back"}
    del x
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


def cpu_oracle_picks(cells, k, picks):
    """The C oracle's greedy picks and fp32 gains on a whole candidate list (checker of mi_parity)."""
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
            "l2": "candidate stream %.0f MB/GPU > 126 MB L2; k-means walks %.1f GB/GPU of resident rows"
                  % (args.mi_candidates * 4 / 1e6, args.km_rows * args.km_d * 4 / 1e9)}


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.perf_counter()
    mi = cpu_mi(args, args.steps, args.warmup)        # one step = one full scan of the bounded sample
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
    if args.skip_mi:
        km = run_kmeans(args, dist, rank, world)
        if rank == 0:
            print(json.dumps({"kmeans": km}), flush=True)
        return
    mi, e2e = run_mi(args, dist, rank, world)
    km = None if args.skip_kmeans else run_kmeans(args, dist, rank, world)
    clocks = sampler.stop()
    if rank == 0:
        pk = peaks()
        cpu = None
        if world == 1 and not args.skip_cpu_baseline:
            cpu = cpu_mi(args, 2000, 3)                    # ~10 s of host work on 16 threads (bounded sample)
            if km is not None:
                km["cpu_baseline"] = cpu_kmeans(args, 200)  # ~5 s
        line = {
            "metric": METRIC, "value": mi["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": mi["ms"] / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(args, world), "us_per_iteration": mi["us_per_iteration"],
            "mi_loop": mi["loop"], "gpu_launches": mi["launches"] + (km["gpu_launches"] if km else 0),
            "e2e": e2e,
            "roofline": {"bound": "hbm", "achieved": mi["achieved_gbs"], "peak": pk["hbm_gbs"], "unit": "GB/s",
                         "frac": mi["achieved_gbs"] / pk["hbm_gbs"],
                         "traffic": 2.125 * args.mi_candidates if mi["loop"].startswith("persistent") else None,
                         "traffic_source": "ncu dram__bytes_read+write per iteration at W = 1e8 (850 MB over a 4-iteration "
                                           "launch), profiles/r01_mi_persist_v3.ncu.txt (2-byte row-partitioned stream + table)",
                         "bytes_per_candidate_accounted": mi["bytes_per_candidate"],
                         "list_order_equivalent_gbs": mi["achieved_gbs"] * 4.0 / mi["bytes_per_candidate"],
                         "kernel": "greedy-MI iteration (gain table + candidate scan + apply), per GPU",
                         "algorithmic_bytes_per_launch": mi["algorithmic_bytes_per_launch"],
                         "peak_source": pk["source"] + " copy bandwidth"},
            "parity_n": (mi["parity_n"] or {}).get("result"), "parity": mi["parity_n"],
            "cell_index_loop": mi.get("cell_index_loop"),
            "cpu_baseline": cpu, "cpu_model": cpu_model(), "kmeans": km, "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if dist:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
