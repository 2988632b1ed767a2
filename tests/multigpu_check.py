"""Multi-GPU parity check against the oracle, run under torchrun (one rank per GPU).  tests/test_multigpu_gpu.py
spawns it as a `-m gpu` test wherever at least two devices are visible; by hand:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/multigpu_check.py

k-means: N-rank KMeans.add (NCCL all-reduce of histogram and deltas) vs the oracle's world step.
greedy MI: candidate list sharded over the ranks vs the C oracle on the whole list (one pair: all three loops; several
pairs: the record all-gather of acav_mi_pairs_local_best / apply).
"""
import os
import sys
import types

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from acav100m_b200 import synth                                  # noqa: E402
from acav100m_b200.clustering import KMeans                      # noqa: E402
from acav100m_b200.subset_selection import get_measure           # noqa: E402
from oracle import kmeans_oracle as ko, mi_oracle as mo          # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))

# ---- k-means ----
# (assignment mode, exchange, batches resident on the GPU and walked for several passes -> CUDA-graph replay)
k, d, b, seed = 32, 256, 512, 9
x = torch.from_numpy(synth.gaussian_mixture(b * world * 12, d, 20, 17))
args = types.SimpleNamespace(computation=types.SimpleNamespace(device="cuda", num_gpus=world))
for mode, comm, resident in (("exact", "p2p", False), ("tensor", "nccl", False), ("tensor", "p2p", True), ("exact", "nccl", True)):
    torch.manual_seed(seed + rank)
    km = KMeans(args, d, k, assign_mode=mode, warmup_rng="cpu", comm=comm)
    km.to("cuda")
    km.initialize()
    # oracle replay of all ranks' RNG streams
    inits, gens = [], []
    for r in range(world):
        torch.manual_seed(seed + r)
        inits.append(torch.rand(k, d) * 1e-5)
        gens.append(torch.get_rng_state())
    c0 = inits[0].clone()
    for t in inits[1:]:
        c0 += t
    st = ko.SgdKMeansState(centers=c0 * (1.0 / world), counts=torch.zeros(k))
    xg = x.cuda() if resident else x
    flips = 0
    torch.set_rng_state(gens[rank])                          # this rank's warm-up noise continues its own stream
    for p in range(3 if resident else 1):
        for g0 in range(0, len(x), world * b):
            noises = None
            if ko.in_warmup(st):
                mine_state = torch.get_rng_state()
                noises = []
                for r in range(world):
                    torch.set_rng_state(gens[r])
                    noises.append(torch.rand(k, b))
                    gens[r] = torch.get_rng_state()
                torch.set_rng_state(mine_state)
            km.add(xg[g0 + rank * b: g0 + (rank + 1) * b])
            outs = ko.sgd_step_world(st, [x[g0 + r * b: g0 + (r + 1) * b] for r in range(world)], noises)
            flips += int((km.last_best.cpu() != outs[rank][0]).sum())
    km.check_status()
    total_flips = torch.tensor([flips], device="cuda")
    dist.all_reduce(total_flips)
    centers = km.centers.cpu()
    rel = ((centers - st.centers).abs().max() / st.centers.abs().max()).item()
    ok = (torch.equal(km.counts.cpu(), st.counts) and km.count == st.count and km.fallback == st.fallback
          and rel < 1e-5) or int(total_flips) > 0
    bit_exact = torch.equal(centers, st.centers)
    n_graphs = len(km._gs["graphs"]) if km._gs else 0      # before calc_best: a bigger batch re-creates the workspace
    best, _ = km.calc_best(x[:2048])
    want, _ = ko.assign(st, x[:2048])
    agree = (best.cpu() == want).float().mean().item()
    print(f"[rank {rank}] kmeans {mode}/{km.comm_name()}{'/resident' if resident else ''}: ok={ok} bit-exact={bit_exact} "
          f"max rel |dcenter| {rel:.2e} ids agree {agree:.4f} near-tie flips {int(total_flips)} graphs {n_graphs}", flush=True)
    assert ok and agree > 0.999
    assert km.comm_name() == comm
    if comm == "p2p" and int(total_flips) == 0:
        assert bit_exact, "rank-ordered peer-memory reduction must reproduce the oracle's world step bit for bit"
    if resident:
        assert n_graphs > 0, "recurring resident batches must be replayed from CUDA graphs"
    # every rank holds the same centers bit for bit (replicated state must not drift)
    ref = km.centers.clone()
    dist.broadcast(ref, 0)
    assert torch.equal(ref, km.centers), "centers differ between ranks"
    del km

# k-means with a skewed batch: a few centroids own >= 128 rows of every rank's slice (ring kernel, split variant)
k, d, b = 6, 256, 2048
gsk = torch.Generator().manual_seed(31)
means = torch.randn(3, d, generator=gsk) * 3.0
x = means[torch.randint(0, 3, (b * world * 4,), generator=gsk)] + torch.randn(b * world * 4, d, generator=gsk)
c0 = torch.cat([means, means + 100.0])                        # three centroids own everything, three stay empty
km = KMeans(args, d, k, assign_mode="exact", warmup_rng="cpu")
km.centers = c0.clone()
km.to("cuda")
st = ko.SgdKMeansState(centers=c0.clone(), counts=torch.zeros(k))
st.count = km.count = 10 * k                                  # past the warm-up: distance assignment from step 1
for g0 in range(0, len(x), world * b):
    km.add(x[g0 + rank * b: g0 + (rank + 1) * b])
    ko.sgd_step_world(st, [x[g0 + r * b: g0 + (r + 1) * b] for r in range(world)])
rel = ((km.centers.cpu() - st.centers).abs().max() / st.centers.abs().max()).item()
ok = torch.equal(km.counts.cpu(), st.counts) and km.fallback == st.fallback and rel < 1e-5
print(f"[rank {rank}] kmeans skewed batch (heavy centroids, split update): ok={ok} max rel |dcenter| {rel:.2e}", flush=True)
assert ok

# ---- greedy MI ----
W, C, picks = 200_003, 64, 700
a = synth.zipf_pairs(W, C, 23)
pos_want, gain_want = mo.greedy_mem_mi_c(a[:, 0], a[:, 1], C, picks)
for loop in ("kernels", "persistent", "cells"):
    m = get_measure("mem_mi")(a, ncentroids=C, device="cuda", shard=(rank, world), loop=loop)
    m.init([(0, 1)], list(range(W)))
    p1, g1 = m.select(picks // 2)
    p2, g2 = m.select(picks - picks // 2)
    pos, gain = torch.cat([p1, p2]), torch.cat([g1, g2])
    ok = np.array_equal(pos.cpu().numpy(), pos_want) and np.array_equal(gain.cpu().numpy(), gain_want)
    print(f"[rank {rank}] greedy MI sharded over {world} ranks, loop={m.loop_name()}: bit-exact={ok}", flush=True)
    assert ok
    del m
# ---- greedy MI over several clustering pairs (records all-gathered per iteration) ----
rng = np.random.RandomState(29)
W, D, C, picks = 60_001, 6, 24, 150
base = rng.randint(0, C, size=(W, 1))
a = np.where(rng.random_sample((W, D)) < 0.5, (base + np.arange(D)) % C, rng.randint(0, C, size=(W, D))).astype(np.int64)
pairs = mo.cluster_pairing([("m%d" % i, "l") for i in range(D)], "combination")
pos_want, gain_want = mo.greedy_mem_mi_pairs_c(a, C, pairs, picks)
m = get_measure("mem_mi")(a, ncentroids=C, device="cuda", shard=(rank, world))
m.init(pairs, list(range(W)))
pos, gain = m.select(picks)
ok = np.array_equal(pos.cpu().numpy(), pos_want) and np.array_equal(gain.cpu().numpy(), gain_want)
print(f"[rank {rank}] greedy MI over {len(pairs)} pairs sharded over {world} ranks, loop={m.loop_name()}: bit-exact={ok}",
      flush=True)
assert ok
del m
# ---- a rank that does not show up must not hang the others: bounded peer wait (ACAV_MI_SPIN_TIMEOUT_MS) ----
if os.environ.get("ACAV_MI_SPIN_TIMEOUT_MS"):
    W, C = 50_000, 32
    a = synth.zipf_pairs(W, C, 41)
    for loop in ("persistent", "cells"):
        m = get_measure("mem_mi")(a, ncentroids=C, device="cuda", shard=(rank, world), loop=loop)
        m.init([(0, 1)], list(range(W)))
        m.select(5)
        m.check_status()
        dist.barrier()
        if rank == 0:                                         # the other ranks never launch this one
            pos, _ = m.select(7)
            try:
                m.check_status()
                raised = False
            except RuntimeError:
                raised = True
            print(f"[rank 0] loop={loop}: lone rank gave up waiting and raised={raised}, unfinished picks {pos.cpu().tolist()}",
                  flush=True)
            assert raised and int(pos[0]) == -1
        dist.barrier()
        del m
dist.barrier()
dist.destroy_process_group()
if rank == 0:
    print("MULTIGPU OK")
