"""world_size-2 gloo runs (CPU) of the multi-GPU host logic in acav100m_b200/parallel.py and of
KMeans.add's distributed orchestration, with the oracle standing in for the CUDA kernels.

What is under test is the PROTOCOL the GPU path uses unchanged: contiguous sharding with global
positions, the 64-bit (score, position) key, "every rank applies the gathered winner", the order of
the two k-means all-reduces, `count` advancing by the global batch, averaging of the inits.
"""
import os
import socket
import types

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from acav100m_b200 import parallel, synth
from oracle import kmeans_oracle as ko, mi_oracle as mo


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _spawn(fn, world, *args):
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_entry, args=(fn, r, world, port, q) + args) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    out = {}
    while not q.empty():
        r, val = q.get()
        out[r] = val
    return [out[r] for r in range(world)]


def _entry(fn, rank, world, port, q, *args):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        q.put((rank, fn(rank, world, *args)))
    finally:
        dist.destroy_process_group()


# ---- key / shard helpers -------------------------------------------------------------------------

def test_key_order_matches_reference_argmax_rule():
    scores = [-3.5, -0.0, 0.0, 1e-12, 0.69314718, 0.6931472, 7.0]
    keys = [parallel.pack_key(s, 10) for s in scores]
    assert keys == sorted(keys) and keys[1] == keys[2]              # -0.0 == +0.0 like torch.max
    assert parallel.pack_key(1.0, 5) > parallel.pack_key(1.0, 6)    # equal score: earliest position wins
    assert parallel.pack_key(1.0, 4_000_000_000) > 0                # 0 stays reserved for "nothing left"
    for s in scores[2:]:
        got, pos = parallel.unpack_key(parallel.pack_key(s, 123456))
        assert got == float(np.float32(s)) and pos == 123456
    assert parallel.combine_pairs([(0, 0), (keys[3], 7), (keys[6], 9), (keys[5], 1)]) == (keys[6], 9)
    assert parallel.unpack_cell(parallel.pack_cell(1023, 17)) == (1023, 17)


def test_shard_bounds_cover_list_in_order():
    for n in (0, 1, 7, 100, 1001):
        for world in (1, 2, 3, 8):
            b = [parallel.shard_bounds(n, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))


# ---- greedy MI over sharded candidates ----------------------------------------------------------

class _OracleMiEngine:
    """CPU stand-in for acav_mi_local_best / acav_mi_apply built from oracle/mi_oracle.py."""

    def __init__(self, assignments, C, candidates, lo, hi):
        self.tab = mo.init_table(1, C)
        self.lo = lo
        self.cells = mo.candidate_cells(assignments, [(0, 1)], candidates[lo:hi])    # [w, 1, 2]
        self.alive = torch.ones(hi - lo, dtype=torch.bool)
        self.picks = []

    def local_best(self, out_pair):
        idx = torch.nonzero(self.alive)[:, 0]
        if idx.numel() == 0:
            out_pair.zero_()
            return
        scores, _, _, _ = mo.candidate_scores(self.tab, self.cells[idx])
        score, j = scores.mean(dim=-1).max(dim=0)
        local = int(idx[j])
        c1, c2 = (int(v) for v in self.cells[local, 0])
        key = parallel.pack_key(score.item(), self.lo + local)
        out_pair[0] = np.array(key, dtype=np.uint64).view(np.int64).item()
        out_pair[1] = parallel.pack_cell(c1, c2)

    def apply(self, all_pairs, world, i):
        pairs = [(int(np.array(int(k), dtype=np.int64).view(np.uint64)), int(c)) for k, c in all_pairs.view(-1, 2).tolist()]
        key, cell = parallel.combine_pairs(pairs)
        score, pos = parallel.unpack_key(key)
        c1, c2 = parallel.unpack_cell(cell)
        cand = torch.tensor([[[c1, c2]]])
        _, NlogN, aloga, blogb = mo.candidate_scores(self.tab, cand)
        mo.apply_pick(self.tab, cand[0], NlogN[0], aloga[0], blogb[0])
        if self.lo <= pos < self.lo + len(self.alive):
            self.alive[pos - self.lo] = False
        self.picks.append((pos, score))


def _mi_rank(rank, world, a, C, n_picks):
    cands = list(range(len(a)))
    lo, hi = parallel.shard_bounds(len(cands), rank, world)
    eng = _OracleMiEngine(a, C, cands, lo, hi)
    parallel.sharded_greedy(eng, dist, world, n_picks, lambda: torch.zeros(2, dtype=torch.int64),
                            lambda w: torch.zeros(2 * w, dtype=torch.int64))
    return eng.picks


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_greedy_equals_single_process(world):
    a = synth.zipf_pairs(700, 12, 41)
    a[0] = 11
    n_picks = 150
    want_S, want_G = mo.greedy_mem_mi(a, 12, [(0, 1)], list(range(700)), n_picks + 1, [])
    outs = _spawn(_mi_rank, world, a, 12, n_picks)
    for picks in outs:                                     # every rank holds the same selection
        assert [p for p, _ in picks] == want_S
        assert [g for _, g in picks] == want_G


class _OraclePairsEngine:
    """CPU stand-in for acav_mi_pairs_local_best / acav_mi_pairs_apply (records = key + the winner's d ids)."""

    def __init__(self, assignments, C, pairs, lo, hi):
        self.pairs, self.d = pairs, assignments.shape[1]
        self.tab = mo.init_table(len(pairs), C)
        self.lo = lo
        self.rows = torch.from_numpy(assignments[lo:hi])
        self.cells = mo.candidate_cells(assignments, pairs, range(lo, hi))
        self.alive = torch.ones(hi - lo, dtype=torch.bool)
        self.picks = []

    @staticmethod
    def _i64(words):
        return torch.from_numpy(np.array(words, dtype=np.uint64).view(np.int64).copy())

    def local_best(self, out_rec):
        idx = torch.nonzero(self.alive)[:, 0]
        if idx.numel() == 0:
            out_rec.zero_()
            return
        scores, _, _, _ = mo.candidate_scores(self.tab, self.cells[idx])
        score, j = scores.mean(dim=-1).max(dim=0)
        local = int(idx[j])
        out_rec.copy_(self._i64(parallel.pack_record(parallel.pack_key(score.item(), self.lo + local),
                                                     self.rows[local].tolist())))

    def apply(self, all_recs, world, i):
        words = all_recs.numpy().view(np.uint64).reshape(world, -1).tolist()
        key, ids = parallel.unpack_record(parallel.combine_records(words), self.d)
        score, pos = parallel.unpack_key(key)
        cand = torch.tensor([[[ids[c1], ids[c2]] for c1, c2 in self.pairs]])
        _, NlogN, aloga, blogb = mo.candidate_scores(self.tab, cand)
        mo.apply_pick(self.tab, cand[0], NlogN[0], aloga[0], blogb[0])
        if self.lo <= pos < self.lo + len(self.alive):
            self.alive[pos - self.lo] = False
        self.picks.append((pos, score))


def _pairs_rank(rank, world, a, C, pairs, n_picks):
    lo, hi = parallel.shard_bounds(len(a), rank, world)
    eng = _OraclePairsEngine(a, C, pairs, lo, hi)
    words = parallel.record_words(a.shape[1])
    parallel.sharded_greedy(eng, dist, world, n_picks, lambda: torch.zeros(words, dtype=torch.int64),
                            lambda w: torch.zeros(words * w, dtype=torch.int64))
    return eng.picks


def test_record_layout_round_trip():
    ids = [0, 65534, 17, 1023, 5, 9, 77]
    rec = parallel.pack_record(parallel.pack_key(0.5, 42), ids)
    assert len(rec) == parallel.record_words(7) == 3
    assert parallel.unpack_record(rec, 7) == (parallel.pack_key(0.5, 42), ids)
    assert parallel.combine_records([[0, 1, 2], rec, parallel.pack_record(parallel.pack_key(0.5, 43), ids)]) == rec
    assert parallel.combine_records([[0, 0, 0]]) is None


def test_sharded_pairs_greedy_equals_single_process():
    rng = np.random.RandomState(43)
    a = rng.randint(0, 7, size=(500, 5)).astype(np.int64)
    pairs = mo.cluster_pairing([("m%d" % i, "l") for i in range(5)], "combination")       # P = 10
    n_picks = 60
    want_pos, want_gain = mo.greedy_mem_mi_pairs_c(a, 7, pairs, n_picks)
    for picks in _spawn(_pairs_rank, 2, a, 7, pairs, n_picks):
        assert [p for p, _ in picks] == want_pos.tolist()
        assert np.array_equal(np.array([g for _, g in picks], dtype=np.float32), want_gain)


# ---- k-means distributed step --------------------------------------------------------------------

def _cpu_protocol_kmeans(d, k, world):
    """KMeans (the product class) with its device hooks replaced by oracle math on CPU."""
    from acav100m_b200.clustering import KMeans

    class CpuHookKMeans(KMeans):
        def _device(self):
            return self.centers.device

        def _prep_batch(self, batch):
            return batch.to(torch.float32)

        def _state(self):
            return ko.SgdKMeansState(self.centers, self.counts, self.count, self.lr, self.initial_rounds,
                                     tuple(self.reinit))

        def _assign(self, batch, want_mean):
            best, mean = ko.assign(self._state(), batch)
            return best, torch.tensor([mean])

        def _histogram(self, batch, best):
            self._best = best
            return torch.zeros(self.centers.shape[0]).scatter_add_(0, best, torch.ones(len(batch)))

        def _update_local(self, batch, counts_b_global, lr, deltas=None):
            lr_eff, fell = ko.effective_lr(lr, counts_b_global.max().item())
            self._fallback_base += int(fell)
            self.counts += counts_b_global
            self.centers *= (1. - counts_b_global * lr_eff)[:, None]
            deltas = torch.zeros_like(self.centers)
            deltas.scatter_add_(0, self._best[:, None].expand(-1, self.centers.shape[1]), batch * lr_eff)
            return deltas

        def _apply_deltas(self, deltas):
            self.centers += deltas

    args = types.SimpleNamespace(computation=types.SimpleNamespace(device="cuda", num_gpus=world))
    return CpuHookKMeans(args, d, k)


def _km_rank(rank, world, x, k, b, seed):
    torch.manual_seed(seed + rank)
    km = _cpu_protocol_kmeans(x.shape[1], k, world)
    assert km.is_distributed
    km.initialize()                                        # averages the per-rank random inits
    init = km.centers.clone()
    for g0 in range(0, len(x) - world * b + 1, world * b):
        km.add(x[g0 + rank * b: g0 + (rank + 1) * b])
    return init.numpy(), km.centers.numpy(), km.counts.numpy(), km.count, km.fallback


def test_kmeans_two_rank_step_equals_world_oracle():
    world, k, d, b, seed = 2, 6, 24, 64, 5
    x = torch.from_numpy(synth.gaussian_mixture(1024, d, 5, 3))
    outs = _spawn(_km_rank, world, x, k, b, seed)
    # single-process replay: same per-rank RNG streams (init draw, then warm-up noise per step)
    gens = []
    inits = []
    for r in range(world):
        torch.manual_seed(seed + r)
        inits.append(torch.rand(k, d) * 1e-5)
        gens.append(torch.get_rng_state())
    st = ko.SgdKMeansState(centers=(inits[0] + inits[1]) * (1.0 / world), counts=torch.zeros(k))
    for g0 in range(0, len(x) - world * b + 1, world * b):
        noises = None
        if ko.in_warmup(st):
            noises = []
            for r in range(world):
                torch.set_rng_state(gens[r])
                noises.append(torch.rand(k, b))
                gens[r] = torch.get_rng_state()
        ko.sgd_step_world(st, [x[g0 + r * b: g0 + (r + 1) * b] for r in range(world)], noises)
    for init, centers, counts, count, fallback in outs:
        assert np.array_equal(init, ((inits[0] + inits[1]) * (1.0 / world)).numpy())
        assert np.array_equal(centers, st.centers.numpy())          # 2 ranks: a+b is order-free
        assert np.array_equal(counts, st.counts.numpy())
        assert count == st.count == (len(x) // (world * b)) * world * b
        assert fallback == st.fallback
