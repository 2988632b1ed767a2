"""GPU parity: CUDA k-means operator (through the C ABI) vs the oracle / reference goldens."""
import os

import numpy as np
import pytest
import torch

from acav100m_b200 import synth
from oracle import gen_golden, kmeans_oracle as ko

pytestmark = pytest.mark.gpu

KM = sorted(gen_golden.KMEANS_CASES)


def load(golden_dir, name):
    return dict(np.load(os.path.join(golden_dir, name + ".npz")))


def make_gpu_kmeans(d, k, sequential=False, **kw):
    from acav100m_b200.clustering import KMeans
    km = KMeans(None, d, k, **kw)
    km.sequential = sequential
    km.to("cuda")
    return km


def state_to_gpu(st, **kw):
    from acav100m_b200.clustering import KMeans
    km = KMeans(None, st.centers.shape[1], st.centers.shape[0], **kw)
    km.centers, km.counts, km.count, km.lr = st.centers.clone(), st.counts.clone(), st.count, st.lr
    km.reinit, km.initial_rounds = st.reinit, st.initial_rounds
    km.to("cuda")
    return km


def assert_ids_match_up_to_fp32_ties(got, want, centers, x, underused=None):
    got, want = np.asarray(got), np.asarray(want)
    differ = got != want
    if differ.any():
        _, d1, d2 = ko.assign_truth_f64(centers, x, underused)
        band = ko.fp32_ambiguity_band(centers, x)
        assert np.all((d2 - d1)[differ] <= band[differ]), "id mismatch outside the fp32 near-tie band"
    return int(differ.sum())


@pytest.mark.parametrize("name", KM)
def test_assign_exact_matches_reference_ids(golden_dir, name):
    case, g = gen_golden.KMEANS_CASES[name], load(golden_dir, name)
    x = synth.gaussian_mixture(case["n"], case["d"], case["k_true"], case["seed"])
    st = ko.SgdKMeansState(centers=torch.from_numpy(g["centers"]), counts=torch.from_numpy(g["counts"]),
                           count=int(g["count"]))
    km = state_to_gpu(st, assign_mode="exact")
    best, mean_d = km.calc_best(torch.from_numpy(x))
    assert best.dtype == torch.int64 and best.is_cuda
    n_diff = assert_ids_match_up_to_fp32_ties(best.cpu().numpy(), g["assign_best"], g["centers"], x,
                                              ko.underused_mask(st).numpy())
    assert n_diff <= 2
    assert mean_d == pytest.approx(float(g["assign_mean_dist"]), rel=1e-5)


def _oracle_trajectory(case):
    """The oracle's run of a golden case (bit-identical to the reference's, tests/test_oracle_golden.py), keeping
    per step the assignment and the state it was computed from."""
    x = torch.from_numpy(synth.gaussian_mixture(case["n"], case["d"], case["k_true"], case["seed"]))
    gen_golden.seed_all(case["seed"])
    st = ko.new_state(case["d"], case["k"])
    st.sequential = bool(case.get("sequential", False))
    steps = []
    for epoch in range(case["epochs"]):
        st.lr = ko.epoch_lr(epoch)
        for xb in gen_golden.kmeans_batches(x, case["batch"]):
            before = st.clone()
            best, _ = ko.sgd_step(st, xb)
            steps.append((before, xb, best.clone()))
    return x, st, steps


def _check_trajectory(case, g, mode):
    """Train through KMeans.add; the centers must equal the reference's BIT FOR BIT unless some step assigned a
    row differently -- and that is only accepted for rows inside the fp32 near-tie band of that step."""
    x, st, steps = _oracle_trajectory(case)
    assert np.array_equal(st.centers.numpy(), g["centers"])      # the oracle is the reference
    gen_golden.seed_all(case["seed"])
    km = make_gpu_kmeans(case["d"], case["k"], assign_mode=mode, warmup_rng="cpu",
                         sequential=bool(case.get("sequential", False)))
    assert np.array_equal(km.centers.cpu().numpy(), g["init_centers"])
    dists, flips, i = [], 0, 0
    for epoch in range(case["epochs"]):
        km.lr = ko.epoch_lr(epoch)
        for xb in gen_golden.kmeans_batches(x, case["batch"]):
            dists.append(km.add(xb))
            before, _, want = steps[i]
            got = km.last_best.cpu()
            if flips == 0 and not torch.equal(got, want):          # first divergence: must be an fp32 near-tie
                assert not ko.in_warmup(before)
                flips += assert_ids_match_up_to_fp32_ties(got.numpy(), want.numpy(), before.centers.numpy(),
                                                          xb.numpy(), ko.underused_mask(before).numpy())
            i += 1
    assert km.count == int(g["count"])
    centers = km.centers.cpu().numpy()
    if flips == 0:
        assert km.fallback == int(g["fallback"])
        assert np.array_equal(km.counts.cpu().numpy(), g["counts"])
        assert np.array_equal(centers, g["centers"]), "identical assignments at every step => identical bits"
    else:
        np.testing.assert_allclose(centers, g["centers"], rtol=2e-5, atol=1e-7)
    np.testing.assert_allclose(np.array(dists), g["step_mean_dist"], rtol=1e-5)
    best, _ = km.calc_best(x)
    assert (best.cpu().numpy() == g["assign_best"]).mean() > 0.998
    return flips


@pytest.mark.parametrize("name", KM)
def test_training_trajectory_matches_reference(golden_dir, name):
    """Whole train loop (warm-up noise from the same seeded CPU generator, lr schedule, fallback,
    re-init scaling) through KMeans.add: centers/counts vs the reference's CPU run."""
    _check_trajectory(gen_golden.KMEANS_CASES[name], load(golden_dir, name), "exact")


# the last six cases put >= 128 rows on a centroid (km_update_stream_kernel, the cp.async ring path): ring wrap-around
# (> 128 rows), index chunk wrap-around (> 1024 rows), row counts that are not multiples of the 8-row groups, a
# d % 4 != 0 control that must stay on the scalar kernel, and the skewed early-training shape b = 8192, k = 3
@pytest.mark.parametrize("b,d,k", [(1, 8, 1), (63, 13, 5), (64, 64, 64), (65, 88, 17), (1000, 352, 33),
                                   (4096, 128, 300), (4096, 128, 8), (4099, 132, 8), (4096, 130, 8),
                                   (5003, 256, 3), (8192, 2304, 3), (8192, 2048, 40),
                                   # fused block prefix + lr decision inside the update kernels up to 32768 rows, four kernels + lr kernel beyond
                                   (16384, 64, 700), (16385, 64, 700), (40000, 32, 50), (1025, 16, 2600), (600, 8, 6200)])
def test_update_is_bit_exact_given_assignments(b, d, k):
    rng = np.random.RandomState(b + d + k)
    x = torch.from_numpy((rng.standard_normal((b, d)) * 10 ** rng.uniform(-2, 2, (b, 1))).astype(np.float32))
    best = torch.from_numpy(rng.randint(0, k, size=b).astype(np.int64))
    if k > 2:
        best[best == 1] = 0                                 # an empty centroid and a heavy one
    for lr in (1e-2, 1e-3):
        st = ko.SgdKMeansState(centers=torch.from_numpy(rng.standard_normal((k, d)).astype(np.float32)),
                               counts=torch.from_numpy(rng.randint(0, 50, k).astype(np.float32)),
                               count=12345, lr=lr)
        km = state_to_gpu(st, assign_mode="exact")
        ko.sgd_step(st, x, best=best)
        from acav100m_b200 import _lib
        xg, bg = x.cuda(), best.cuda()
        counts_b = torch.empty(k, dtype=torch.float32, device="cuda")
        ws = km._workspace(b)
        s = _lib.stream_ptr()
        _lib.call("acav_kmeans_histogram", ws, _lib.ptr(bg), b, _lib.ptr(counts_b), s)
        _lib.call("acav_kmeans_update_fused", ws, _lib.ptr(xg), b, xg.stride(0), _lib.ptr(counts_b), float(lr),
                  _lib.ptr(km.centers), _lib.ptr(km.counts), _lib.ptr(km._fallback_dev), s)
        assert np.array_equal(counts_b.cpu().numpy(), np.bincount(best.numpy(), minlength=k).astype(np.float32))
        assert np.array_equal(km.counts.cpu().numpy(), st.counts.numpy())
        assert np.array_equal(km.centers.cpu().numpy(), st.centers.numpy()), "row-order fp32 sum must be bit-exact"
        assert km.fallback == st.fallback


@pytest.mark.parametrize("b,d,k", [(3000, 96, 40), (4096, 128, 8), (5003, 256, 3), (4099, 130, 6)])
def test_split_update_equals_fused(b, d, k):
    """update_local + apply_deltas (the multi-GPU split) reproduces update_fused bit for bit -- below and above
    the 128-rows-per-centroid threshold of the ring kernel -- and the deltas are the oracle's row-order sums."""
    from acav100m_b200 import _lib
    rng = np.random.RandomState(3)
    xc = torch.from_numpy(rng.standard_normal((b, d)).astype(np.float32))
    bc = torch.from_numpy(rng.randint(0, k, size=b).astype(np.int64))
    x, best = xc.cuda(), bc.cuda()
    c0 = torch.from_numpy(rng.standard_normal((k, d)).astype(np.float32))
    outs = []
    for split in (False, True):
        st = ko.SgdKMeansState(centers=c0.clone(), counts=torch.zeros(k), count=999)
        km = state_to_gpu(st)
        ws = km._workspace(b)
        s = _lib.stream_ptr()
        counts_b = torch.empty(k, dtype=torch.float32, device="cuda")
        _lib.call("acav_kmeans_histogram", ws, _lib.ptr(best), b, _lib.ptr(counts_b), s)
        if split:
            deltas = torch.empty(k, d, dtype=torch.float32, device="cuda")
            _lib.call("acav_kmeans_update_local", ws, _lib.ptr(x), b, d, _lib.ptr(counts_b), 0.01,
                      _lib.ptr(km.centers), _lib.ptr(km.counts), _lib.ptr(deltas), None, s)
            lr_eff = ko.effective_lr(0.01, float(counts_b.max().item()))[0]
            want = ko.sequential_scatter_sum((xc * lr_eff).numpy(), bc.numpy(), k)
            assert np.array_equal(deltas.cpu().numpy(), want), "per-rank deltas must be strict row-order sums"
            _lib.call("acav_kmeans_apply_deltas", _lib.ptr(km.centers), _lib.ptr(deltas), k * d, s)
        else:
            _lib.call("acav_kmeans_update_fused", ws, _lib.ptr(x), b, d, _lib.ptr(counts_b), 0.01,
                      _lib.ptr(km.centers), _lib.ptr(km.counts), None, s)
        outs.append(km.centers.cpu().numpy())
    assert np.array_equal(outs[0], outs[1])
    st = ko.SgdKMeansState(centers=c0.clone(), counts=torch.zeros(k), count=999)
    ko.sgd_step(st, xc, best=bc)
    assert np.array_equal(outs[0], st.centers.numpy())


def test_update_requires_histogram_first():
    from acav100m_b200 import _lib
    km = make_gpu_kmeans(8, 4)
    ws = km._workspace(16)
    x = torch.zeros(16, 8, device="cuda")
    cb = torch.zeros(4, device="cuda")
    with pytest.raises(_lib.AcavError) as e:
        _lib.call("acav_kmeans_update_fused", ws, _lib.ptr(x), 16, 8, _lib.ptr(cb), 0.01, _lib.ptr(km.centers),
                  _lib.ptr(km.counts), None, _lib.stream_ptr())
    assert e.value.status == -3


@pytest.mark.parametrize("k,b", [(1, 1), (16, 32), (37, 1000), (256, 4097)])
def test_warmup_noise_assign_is_first_argmin(k, b):
    from acav100m_b200 import _lib
    g = torch.Generator().manual_seed(k * b)
    noise = torch.rand(k, b, generator=g)
    noise[:, ::7] = noise[0, ::7].clone()                   # whole-column ties -> index 0 must win
    want_d, want_i = noise.min(axis=0)
    ng = noise.cuda()
    best = torch.empty(b, dtype=torch.int64, device="cuda")
    mind = torch.empty(b, dtype=torch.float32, device="cuda")
    mean = torch.empty(1, dtype=torch.float32, device="cuda")
    _lib.call("acav_kmeans_assign_noise", _lib.ptr(ng), k, b, _lib.ptr(best), _lib.ptr(mind), _lib.ptr(mean),
              _lib.stream_ptr())
    assert torch.equal(best.cpu(), want_i)
    assert torch.equal(mind.cpu(), want_d)
    assert mean.item() == pytest.approx(want_d.mean().item(), rel=1e-6)


@pytest.mark.parametrize("b,d,k", [(1, 3, 2), (5, 7, 3), (130, 70, 129), (513, 2304, 20)])
def test_assign_exact_ragged_shapes_and_ties(b, d, k):
    rng = np.random.RandomState(b * 31 + d)
    c = rng.standard_normal((k, d)).astype(np.float32)
    c[k - 1] = c[0]                                          # duplicate centroid: lower index wins
    x = (c[rng.randint(0, k, size=b)] + 0.01 * rng.standard_normal((b, d))).astype(np.float32)
    counts = rng.randint(0, 100, k).astype(np.float32)
    st = ko.SgdKMeansState(centers=torch.from_numpy(c), counts=torch.from_numpy(counts), count=40 * k)
    km = state_to_gpu(st, assign_mode="exact")
    best, mean_d = km.calc_best(torch.from_numpy(x))
    under = ko.underused_mask(st).numpy()
    want, d1, _ = ko.assign_truth_f64(c, x, under)
    assert_ids_match_up_to_fp32_ties(best.cpu().numpy(), want, c, x, under)
    assert (best.cpu().numpy() != k - 1).all() or under[k - 1] != under[0]
    scale = np.abs(d1).mean() + (x.astype(np.float64) ** 2).sum(1).mean()
    assert abs(mean_d - d1.mean()) <= 1e-5 * scale


def test_assign_strided_batch_and_empty_batch():
    rng = np.random.RandomState(9)
    big = torch.from_numpy(rng.standard_normal((200, 96)).astype(np.float32)).cuda()
    view = big[:, :64]                                       # row stride 96, d = 64
    km = make_gpu_kmeans(64, 8, assign_mode="exact")
    km.count = 10_000
    km.centers = torch.from_numpy(rng.standard_normal((8, 64)).astype(np.float32)).cuda()
    b1, _ = km.calc_best(view)
    b2, _ = km.calc_best(view.contiguous())
    assert torch.equal(b1, b2)
    best, mean = km.calc_best(torch.zeros(0, 64))
    assert best.numel() == 0 and np.isnan(mean)


# ---- tcgen05 tensor-core assignment ---------------------------------------------------------------

def _assign_both_modes(x, centers, counts, count, tile_variant=0):
    from acav100m_b200 import _lib
    outs = {}
    b, d = x.shape
    k = centers.shape[0]
    for mode in ("exact", "tensor"):
        st = ko.SgdKMeansState(centers=centers.clone(), counts=counts.clone(), count=count)
        km = state_to_gpu(st, assign_mode=mode, tile_variant=tile_variant)
        ws = km._workspace(b)
        xg = x.cuda()
        best = torch.empty(b, dtype=torch.int64, device="cuda")
        mind = torch.empty(b, dtype=torch.float32, device="cuda")
        mean = torch.empty(1, dtype=torch.float32, device="cuda")
        nref = torch.zeros(2, dtype=torch.int32, device="cuda")
        _lib.call("acav_kmeans_assign", ws, _lib.ptr(xg), b, d, _lib.ptr(km.centers), _lib.ptr(km.counts),
                  km.underused_threshold(), float(km.reinit[1]), _lib.ptr(best), _lib.ptr(mind), _lib.ptr(mean),
                  _lib.ptr(nref), km._mode(), _lib.stream_ptr())
        outs[mode] = (best.cpu().numpy(), mind.cpu().numpy(), mean.item(), nref.cpu().tolist())
    return outs


@pytest.mark.parametrize("b,d,k,clustered", [
    (128, 64, 16, True), (1, 64, 16, True), (129, 64, 256, True), (1000, 128, 256, True), (300, 88, 13, True),
    (4097, 512, 300, True), (8192, 2048, 1024, True), (20000, 704, 1024, True), (3000, 2304, 32, True),
    (2048, 256, 512, False), (5000, 128, 1500, False)])
@pytest.mark.parametrize("tile_variant", [1, 2, 3])
def test_assign_tensor_equals_exact(b, d, k, clustered, tile_variant):
    """Screen + re-check must reproduce the exact kernel's ids on EVERY row (that is the design claim),
    and the distances handed back must be the exact ones -- for each of the three tcgen05 tile shapes
    (one CTA 128x256, CTA pair 256x256, CTA pair 256x512)."""
    if clustered:
        x = torch.from_numpy(synth.gaussian_mixture(b, d, max(k // 2, 2), b + d))
        c = torch.from_numpy(synth.gaussian_mixture(k, d, max(k // 2, 2), b + d))
    else:
        g = torch.Generator().manual_seed(b + k)
        x, c = torch.randn(b, d, generator=g), torch.randn(k, d, generator=g)
    counts = torch.full((k,), 100.0)
    counts[::3] = 0.0                                          # a third of the centroids get the /5 scaling
    outs = _assign_both_modes(x, c, counts, 50 * k, tile_variant)
    be, me, mean_e, _ = outs["exact"]
    bt, mt, mean_t, nref = outs["tensor"]
    assert np.array_equal(be, bt), "%d rows differ" % int((be != bt).sum())
    np.testing.assert_allclose(mt, me, rtol=1e-6, atol=1e-6 * np.abs(me).max())
    assert mean_t == pytest.approx(mean_e, rel=1e-6)
    assert nref[0] + nref[1] <= b


@pytest.mark.parametrize("state", ["converged", "early"])
def test_assign_vs_fp64_truth_at_baseline_shape(state):
    """BASELINE config 3's step shape (8192 rows x 2048, K = 1024): the exact kernel AND the tcgen05 path against
    the fp64 evaluation of the reference formula (sgd_clustering.py:72-78) -- ids equal except rows whose two best
    centroids are closer than the fp32 rounding band, and there the CUDA paths still agree with each other."""
    b, d, k = 8192, 2048, 1024
    g = torch.Generator().manual_seed(11)
    means = torch.randn(k, d, generator=g) * 3.0
    x = means[torch.randint(0, k, (b,), generator=g)] + torch.randn(b, d, generator=g)
    if state == "converged":
        c = means + 0.02 * torch.randn(k, d, generator=g)
        counts = torch.full((k,), 500.0)
        counts[::5] = 0.0
    else:                                                       # early training: centroids still near the origin
        c = torch.rand(k, d, generator=g) * 1e-5 + 0.01 * means * (torch.rand(k, 1, generator=g) < 0.02)
        counts = torch.zeros(k)
        counts[:8] = 3000.0
    count = 40 * k
    outs = _assign_both_modes(x, c, counts, count)
    st = ko.SgdKMeansState(centers=c, counts=counts, count=count)
    under = ko.underused_mask(st).numpy()
    want, d1, _ = ko.assign_truth_f64(c.numpy(), x.numpy(), under)
    for mode in ("exact", "tensor"):
        n_diff = assert_ids_match_up_to_fp32_ties(outs[mode][0], want, c.numpy(), x.numpy(), under)
        assert n_diff <= (8 if state == "converged" else b)
    assert np.array_equal(outs["exact"][0], outs["tensor"][0])
    scale = np.abs(d1) + (x.numpy().astype(np.float64) ** 2).sum(1)
    assert np.all(np.abs(outs["exact"][1] - d1) <= 4e-6 * scale)


@pytest.mark.parametrize("name", KM)
def test_training_trajectory_tensor_mode_matches_reference(golden_dir, name):
    _check_trajectory(gen_golden.KMEANS_CASES[name], load(golden_dir, name), "tensor")


def test_assign_all_overlapped_pass_equals_chunked_calc_best():
    n, d, k = 40_000, 256, 300
    g = torch.Generator().manual_seed(4)
    means = torch.randn(k, d, generator=g) * 3
    x = means[torch.randint(0, k, (n,), generator=g)] + torch.randn(n, d, generator=g)
    st = ko.SgdKMeansState(centers=means + 0.05 * torch.randn(k, d, generator=g), counts=torch.full((k,), 50.0),
                           count=20 * k)
    st.counts[::4] = 0.0
    km = state_to_gpu(st, assign_mode="tensor")
    xg = x.cuda()
    want = torch.cat([km.calc_best(xg[lo:lo + 7000])[0] for lo in range(0, n, 7000)])
    got = km.assign_all(xg, chunk=6000)                      # ragged last chunk, 7 chunks over 2 workspaces
    assert torch.equal(got, want)
    best, mean = km.calc_best(xg[:5000], distance=False)
    assert mean is None and torch.equal(best, want[:5000])
    assert km.add(xg[:4096], distance=False) is None
    km_exact = state_to_gpu(st, assign_mode="exact")
    assert torch.equal(km_exact.assign_all(xg, chunk=9000), want)


@pytest.mark.parametrize("mode,distance", [("exact", True), ("tensor", False), ("tensor", True)])
def test_graph_replayed_steps_equal_eager_steps_bit_for_bit(mode, distance):
    """KMeans.add replays the steady-state step from a CUDA graph cached per batch address (graph='auto').  Same
    kernels, same arguments: centers, counts, count, fallback and the assignments must equal the eager path's bit for
    bit over several epochs (graphs captured in epoch 0 are replayed later; the lr change opens new graphs)."""
    n, d, k, b = 6 * 1024, 192, 40, 1024
    x = torch.from_numpy(synth.gaussian_mixture(n, d, 24, 5)).cuda()
    runs = []
    for graph in (False, "auto"):
        torch.manual_seed(77)                                   # same init and the same warm-up noise stream for both
        km = make_gpu_kmeans(d, k, assign_mode=mode, warmup_rng="cpu", graph=graph)
        log = []
        for epoch in range(4):
            km.lr = 1e-2 if epoch < 2 else 1e-3
            for lo in range(0, n, b):
                out = km.add(x[lo:lo + b], sync=True, distance=distance)
                log.append((out, km.last_best.clone()))
            log.append((km.count, km.fallback, km.counts.clone(), km.centers.clone()))
        runs.append((km, log))
    (eager, log_e), (graphed, log_g) = runs
    assert len(log_e) == len(log_g)
    for i, (a, g) in enumerate(zip(log_e, log_g)):
        assert a[:-2] == g[:-2] if len(a) == 4 else a[0] == g[0], "entry %d" % i
        assert all(torch.equal(ta, tg) for ta, tg in zip(a[-2:] if len(a) == 4 else a[1:], g[-2:] if len(g) == 4 else g[1:])), \
            "entry %d differs between the eager and the graph-replayed run" % i
    gs = graphed._gs
    assert gs is not None and len(gs["graphs"]) >= 6, "steady-state steps must have been captured"      # 6 batches x 2 lr
    assert eager._gs is None
    # a state that pickles (checkpoint path) and keeps training
    import pickle
    km2 = pickle.loads(pickle.dumps(graphed))
    km2.to("cuda")
    a = km2.add(x[:b]); e = eager.add(x[:b])
    assert a == e and torch.equal(km2.centers, eager.centers)


def _world_step_forced(st, xs, bests, lr):
    """oracle/kmeans_oracle.py::sgd_step_world with the assignments given (rank-ordered sums)."""
    k, d = st.centers.shape
    counts = torch.zeros(k)
    for best, xb in zip(bests, xs):
        counts += torch.zeros(k).scatter_add_(0, best, torch.ones(len(xb)))
    lr_eff, fell = ko.effective_lr(lr, counts.max().item())
    st.fallback += int(fell)
    st.counts += counts
    st.centers *= (1. - counts * lr_eff)[:, None]
    deltas = torch.zeros_like(st.centers)
    for best, xb in zip(bests, xs):
        local = torch.zeros_like(st.centers)
        local.scatter_add_(0, best[:, None].expand(-1, d), xb * lr_eff)
        deltas += local
    st.centers = st.centers + deltas


@pytest.mark.parametrize("world,b,d,k", [(1, 3000, 96, 40), (1, 4096, 128, 8), (2, 1000, 64, 7), (2, 4096, 256, 6),
                                         (3, 700, 128, 16), (4, 2048, 512, 64)])
def test_peer_memory_step_is_the_rank_ordered_world_step(world, b, d, k, monkeypatch):
    """acav_kmeans_update_p2p (histogram exchange, deltas pushed into the owner's buffer by the update kernels,
    rank-ordered owner-side sum, row broadcast) with `world` ranks driven from this process on one device, one stream
    each: all ranks end with the SAME bits, equal to the oracle's rank-ordered world step -- which for world = 1 is
    update_fused.  Covers centroids above and below the 128-row ring-kernel threshold and k % world != 0."""
    from acav100m_b200 import _lib
    monkeypatch.setenv("ACAV_KM_SPIN_TIMEOUT_MS", "5000")
    rng = np.random.RandomState(world * 1000 + b)
    c0 = torch.from_numpy(rng.standard_normal((k, d)).astype(np.float32))
    st = ko.SgdKMeansState(centers=c0.clone(), counts=torch.from_numpy(rng.randint(0, 30, k).astype(np.float32)), count=7777)
    ranks = []
    for r in range(world):
        km = state_to_gpu(st)
        comm = _lib.c_vp()
        _lib.call("acav_kmeans_comm_create", _lib.ctypes.byref(comm), k, d, world, r)
        mine = (_lib.ctypes.c_ubyte * _lib.load().acav_kmeans_comm_handle_bytes())()
        _lib.call("acav_kmeans_comm_export", comm, mine)
        ranks.append(dict(km=km, comm=comm, ws=km._workspace(b), stream=torch.cuda.Stream(),
                          fb=torch.zeros(1, dtype=torch.int32, device="cuda"),
                          counts_b=torch.empty(k, dtype=torch.float32, device="cuda")))
    if world > 1:
        arenas = (_lib.c_vp * world)(*[_lib.load().acav_kmeans_comm_arena(R["comm"]) for R in ranks])
        for R in ranks:
            _lib.call("acav_kmeans_comm_connect_ptrs", R["comm"], arenas)
    try:
        for step, lr in enumerate((1e-2, 1e-3, 1e-2)):
            xs = [torch.from_numpy((rng.standard_normal((b, d)) * 10 ** rng.uniform(-1, 1, (b, 1))).astype(np.float32))
                  for _ in range(world)]
            bests = [torch.from_numpy(rng.randint(0, k, size=b).astype(np.int64)) for _ in range(world)]
            for bb in bests:
                bb[bb == 1] = 0                                  # an empty centroid and a heavy one
            xg = [t.cuda() for t in xs]
            bg = [t.cuda() for t in bests]
            torch.cuda.synchronize()
            for r, R in enumerate(ranks):
                with torch.cuda.stream(R["stream"]):
                    s = _lib.stream_ptr()
                    _lib.call("acav_kmeans_histogram", R["ws"], _lib.ptr(bg[r]), b, _lib.ptr(R["counts_b"]), s)
                    _lib.call("acav_kmeans_update_p2p", R["ws"], R["comm"], _lib.ptr(xg[r]), b, d, _lib.ptr(R["counts_b"]),
                              float(lr), _lib.ptr(R["km"].centers), _lib.ptr(R["km"].counts), _lib.ptr(R["fb"]), s)
            torch.cuda.synchronize()
            _world_step_forced(st, xs, bests, lr)
            for r, R in enumerate(ranks):
                status = _lib.ctypes.c_int32(7)
                _lib.call("acav_kmeans_comm_status", R["comm"], _lib.ctypes.byref(status), _lib.stream_ptr())
                assert status.value == 0
                assert np.array_equal(R["km"].counts.cpu().numpy(), st.counts.numpy()), (step, r)
                assert np.array_equal(R["km"].centers.cpu().numpy(), st.centers.numpy()), \
                    "step %d rank %d: not the rank-ordered sum" % (step, r)
                assert int(R["fb"].item()) == st.fallback
    finally:
        torch.cuda.synchronize()
        for R in ranks:
            _lib.load().acav_kmeans_comm_destroy(R["comm"])


def test_peer_memory_step_gives_up_on_a_missing_rank(monkeypatch):
    """A rank whose peer never runs the step must come back with the status word set, not hang."""
    from acav100m_b200 import _lib
    monkeypatch.setenv("ACAV_KM_SPIN_TIMEOUT_MS", "300")
    k, d, b = 8, 64, 256
    rng = np.random.RandomState(1)
    st = ko.SgdKMeansState(centers=torch.from_numpy(rng.standard_normal((k, d)).astype(np.float32)), counts=torch.zeros(k), count=99)
    comms, kms = [], []
    for r in range(2):
        kms.append(state_to_gpu(st))
        comm = _lib.c_vp()
        _lib.call("acav_kmeans_comm_create", _lib.ctypes.byref(comm), k, d, 2, r)
        mine = (_lib.ctypes.c_ubyte * _lib.load().acav_kmeans_comm_handle_bytes())()
        _lib.call("acav_kmeans_comm_export", comm, mine)
        comms.append(comm)
    arenas = (_lib.c_vp * 2)(*[_lib.load().acav_kmeans_comm_arena(c) for c in comms])
    for c in comms:
        _lib.call("acav_kmeans_comm_connect_ptrs", c, arenas)
    x = torch.from_numpy(rng.standard_normal((b, d)).astype(np.float32)).cuda()
    best = torch.from_numpy(rng.randint(0, k, size=b).astype(np.int64)).cuda()
    counts_b = torch.empty(k, dtype=torch.float32, device="cuda")
    ws, s = kms[0]._workspace(b), _lib.stream_ptr()
    _lib.call("acav_kmeans_histogram", ws, _lib.ptr(best), b, _lib.ptr(counts_b), s)
    _lib.call("acav_kmeans_update_p2p", ws, comms[0], _lib.ptr(x), b, d, _lib.ptr(counts_b), 0.01,
              _lib.ptr(kms[0].centers), _lib.ptr(kms[0].counts), None, s)            # rank 1 never shows up
    status = _lib.ctypes.c_int32(0)
    _lib.call("acav_kmeans_comm_status", comms[0], _lib.ctypes.byref(status), s)
    assert status.value == 1
    for c in comms:
        _lib.load().acav_kmeans_comm_destroy(c)


@pytest.mark.parametrize("b,d,k", [(64, 13, 5), (1000, 352, 33), (4096, 128, 8), (5003, 256, 3)])
def test_sequential_update_is_bit_exact_given_assignments(b, d, k):
    """`sequential=True` (sgd_clustering.py:103-109): one row at a time in batch order -- per centroid a row-ordered
    recurrence, run by the same kernels as the fast update (light, scalar and TMA heavy-centroid paths)."""
    from acav100m_b200 import _lib
    rng = np.random.RandomState(b + d)
    x = torch.from_numpy(rng.standard_normal((b, d)).astype(np.float32))
    best = torch.from_numpy(rng.randint(0, k, size=b).astype(np.int64))
    if k > 2:
        best[best == 1] = 0
    st = ko.SgdKMeansState(centers=torch.from_numpy(rng.standard_normal((k, d)).astype(np.float32)),
                           counts=torch.from_numpy(rng.randint(0, 50, k).astype(np.float32)), count=999, lr=1e-2,
                           sequential=True)
    km = state_to_gpu(st)
    ko.sgd_step(st, x, best=best)
    xg, bg = x.cuda(), best.cuda()
    counts_b = torch.empty(k, dtype=torch.float32, device="cuda")
    ws, s = km._workspace(b), _lib.stream_ptr()
    _lib.call("acav_kmeans_histogram", ws, _lib.ptr(bg), b, _lib.ptr(counts_b), s)
    _lib.call("acav_kmeans_update_sequential", ws, _lib.ptr(xg), b, d, _lib.ptr(counts_b), 1e-2, _lib.ptr(km.centers),
              _lib.ptr(km.counts), s)
    assert np.array_equal(km.counts.cpu().numpy(), st.counts.numpy())
    assert np.array_equal(km.centers.cpu().numpy(), st.centers.numpy())
