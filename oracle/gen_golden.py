"""Generate tests/golden/*.npz by running the UNMODIFIED reference on CPU (this container only).

TEST INFRASTRUCTURE ONLY.  Usage:  python oracle/gen_golden.py [case ...]   (needs /root/reference mounted).
The reference never seeds its RNGs (SURVEY.md headline fact 4); this harness seeds torch / random /
numpy itself and records the seed with every fixture.  Inputs are regenerated from the seed by
``acav100m_b200.synth`` (numpy legacy RandomState, frozen stream); a float64 checksum of the inputs
is stored so a drifting generator is detected rather than silently compared.
"""
import os
import random
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from acav100m_b200 import synth                     # noqa: E402
from oracle import ref_shims                        # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")

KMEANS_CASES = {
    # name: n, d, k, k_true, batch, epochs, seed          (c1 = BASELINE config 1, reference defaults)
    "kmeans_c1_audio": dict(n=1000, d=128, k=16, k_true=16, batch=32, epochs=2, seed=1001),
    "kmeans_c1_visual": dict(n=1000, d=128, k=16, k_true=16, batch=32, epochs=2, seed=2001),
    # large batch: lr fallback (sgd_clustering.py:116-119) and under-used re-init scaling (:76-77)
    "kmeans_bigbatch": dict(n=4096, d=64, k=32, k_true=20, batch=1024, epochs=3, seed=1002),
    # ragged feature dim / k not a power of two, epoch 5+ changes lr (run_clustering.py:168)
    "kmeans_ragged": dict(n=777, d=88, k=13, k_true=9, batch=37, epochs=6, seed=1003),
    # the slow sequential update branch (sgd_clustering.py:103-109; `sequential` is an attribute, off by default :32)
    "kmeans_sequential": dict(n=900, d=36, k=7, k_true=5, batch=150, epochs=2, seed=1004, sequential=True),
}

MI_CASES = {
    # name: v, c, dcols, subset, pairing, shuffle, seed
    "mi_c1": dict(v=1000, c=16, dcols=2, subset=200, pairing="combination", shuffle=False, seed=1001),
    "mi_c20_shuffled": dict(v=600, c=20, dcols=2, subset=150, pairing="combination", shuffle=True, seed=1004),
    "mi_p3": dict(v=400, c=8, dcols=3, subset=80, pairing="combination", shuffle=False, seed=1005),
    "mi_dense_small": dict(v=300, c=6, dcols=2, subset=299, pairing="combination", shuffle=False, seed=1006),
    # P = 10 and the reference default P = 45 (ten clusterings, `combination`): the mean over pairs goes through
    # torch's vectorised inner-dimension sum (P >= 8), whose addition order the C oracle and the device restate
    "mi_p10": dict(v=300, c=6, dcols=5, subset=60, pairing="combination", shuffle=False, seed=1008),
    "mi_p45": dict(v=250, c=5, dcols=10, subset=50, pairing="combination", shuffle=True, seed=1009),
}


AMI_CASES = {
    # adjusted MI (measures/mi.py:212-262): dense toy-size measure, reachable as --measure_name=ami
    "ami_small": dict(v=200, c=5, dcols=2, subset=60, pairing="combination", shuffle=False, seed=1010),
    "ami_p3": dict(v=150, c=4, dcols=3, subset=40, pairing="combination", shuffle=False, seed=1011),
}

BATCH_MI_CASES = {
    # the reference CLI's default measure (config.py:45): B = 20 candidates per iteration, keep top k = 4
    "bmi_p3": dict(v=400, c=8, dcols=3, subset=80, seed=1005, keep_unselected=True),
    "bmi_p1_c20": dict(v=600, c=20, dcols=2, subset=120, seed=1004, keep_unselected=True),
    "bmi_drop_unselected": dict(v=500, c=6, dcols=2, subset=60, seed=1007, keep_unselected=False),
}


def seed_all(seed):
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)


def kmeans_batches(x, batch):
    """drop_last=True, in-order mini-batches (data/clustering.py:17-33 with shuffle ignored :29)."""
    n = (len(x) // batch) * batch
    return [x[i:i + batch] for i in range(0, n, batch)]


def run_reference_kmeans(case):
    KMeans = ref_shims.load_reference_kmeans()
    x = torch.from_numpy(synth.gaussian_mixture(case["n"], case["d"], case["k_true"], case["seed"]))
    seed_all(case["seed"])
    km = KMeans(ref_shims.reference_kmeans_args(), case["d"], case["k"])
    km.sequential = bool(case.get("sequential", False))
    init_centers = km.centers.clone()
    dists, trace_best = [], []
    with ref_shims.cuda_is_identity(), torch.no_grad():
        for epoch in range(case["epochs"]):
            km.lr = 0.1 ** (2 + epoch // 5)                       # run_clustering.py:168
            for xb in kmeans_batches(x, case["batch"]):
                dists.append(km.add(xb))
        best, mean_d = km.calc_best(x)                            # assign pass, one batch
    return dict(
        seed=case["seed"], x_checksum=float(x.double().sum()),
        init_centers=init_centers.numpy(), centers=km.centers.numpy(), counts=km.counts.numpy(),
        count=km.count, fallback=km.fallback, step_mean_dist=np.array(dists, dtype=np.float64),
        assign_best=best.numpy(), assign_mean_dist=mean_d,
    )


def write_reference_checkpoint(path):
    """A ver1 centroid checkpoint exactly as the reference's default flags write it (run_clustering.py:110-116 with
    `save_scheme_ver2` unset): ``torch.save`` of ``{model: {layer: <sgd_clustering.KMeans object>}}`` -- the objects of
    the UNMODIFIED reference class after a few training steps.  tests/test_host_formats.py reads it back through
    acav100m_b200.clustering.checkpoint without the reference on sys.path."""
    KMeans = ref_shims.load_reference_kmeans()
    seed_all(21)
    tree = {}
    expect = {}
    with ref_shims.cuda_is_identity(), torch.no_grad():
        for model, dims in (("layer_vggish", (8, 12)), ("layer_slow_fast", (16,))):
            tree[model] = {}
            for i, d in enumerate(dims):
                km = KMeans(ref_shims.reference_kmeans_args(), d, 4)
                x = torch.from_numpy(synth.gaussian_mixture(96, d, 3, 21 + i))
                for xb in kmeans_batches(x, 16):
                    km.add(xb)
                km.args = None                                   # the reference pickles its munch tree here; not needed
                tree[model]["layer_%d" % i] = km
                expect["%s/layer_%d/centers" % (model, i)] = km.centers.numpy().copy()
                expect["%s/layer_%d/counts" % (model, i)] = km.counts.numpy().copy()
                expect["%s/layer_%d/count" % (model, i)] = np.int64(km.count)
                expect["%s/layer_%d/fallback" % (model, i)] = np.int64(km.fallback)
    torch.save(tree, path)
    np.savez_compressed(path + ".expect.npz", **expect)


def mi_assignments(case):
    if case["dcols"] == 2:
        return synth.zipf_pairs(case["v"], case["c"], case["seed"])
    rng = np.random.RandomState(case["seed"])
    base = rng.randint(0, case["c"], size=(case["v"], 1))
    noise = rng.randint(0, case["c"], size=(case["v"], case["dcols"]))
    keep = rng.random_sample((case["v"], case["dcols"])) < 0.6
    a = np.where(keep, (base + np.arange(case["dcols"])) % case["c"], noise).astype(np.int64)
    a[0, :] = case["c"] - 1                                        # make sure max()+1 == c
    return a


def run_reference_mi(case, measure_name):
    get_measure, get_pairing = ref_shims.load_reference_measures()
    a = mi_assignments(case)
    assert a.max() + 1 == case["c"], (a.max(), case["c"])
    seed_all(case["seed"])
    keys = [("m%d" % i, "layer") for i in range(case["dcols"])]
    pairs = get_pairing(keys, case["pairing"])
    v = a.shape[0]
    bsz = min(20, v - 1)
    measure = get_measure(measure_name)(a, ncentroids=case["c"], batch_size=bsz,
                                        selection_size=min(4, bsz), device="cpu", keep_unselected=True)
    candidates = list(range(v))
    if case["shuffle"]:
        random.shuffle(candidates)                                  # run_greedy.py:37-40
    order = np.array(candidates, dtype=np.int64)
    start, candidates = [candidates[0]], candidates[1:]
    measure.init(pairs, candidates)
    S, GAIN, _, _ = measure.run_greedy(case["subset"], start, None, verbose=False, log_every=1,
                                       log_times=None, node_rank=None, pid=None)
    return dict(seed=case["seed"], assignments=a.astype(np.int16), candidate_order=order,
                pairs=np.array(pairs, dtype=np.int64), S=np.array(S, dtype=np.int64),
                GAIN=np.array(GAIN, dtype=np.float64), subset=case["subset"], c=case["c"])


def run_reference_batch_mi(case):
    import contextlib
    import io
    get_measure, get_pairing = ref_shims.load_reference_measures()
    a = mi_assignments(case)
    keys = [("m%d" % i, "layer") for i in range(case["dcols"])]
    pairs = get_pairing(keys, "combination")
    v = a.shape[0]
    seed_all(case["seed"])
    measure = get_measure("batch_mi")(a, ncentroids=case["c"], batch_size=min(20, v - 1), selection_size=4,
                                      device="cpu", keep_unselected=case["keep_unselected"])
    candidates = list(range(v))
    start, candidates = [candidates[0]], candidates[1:]
    measure.init(pairs, candidates)
    with contextlib.redirect_stdout(io.StringIO()):
        S, GAIN, _, _ = measure.run_greedy(case["subset"], start, None, verbose=False, log_every=1,
                                           log_times=None, node_rank=None, pid=None)
    return dict(seed=case["seed"], assignments=a.astype(np.int16), pairs=np.array(pairs, dtype=np.int64),
                S=np.array(S, dtype=np.int64), GAIN=np.array(GAIN, dtype=np.float64), subset=case["subset"],
                c=case["c"], keep_unselected=case["keep_unselected"])


def main():
    assert ref_shims.reference_available(), "needs /root/reference"
    os.makedirs(GOLDEN, exist_ok=True)
    only = set(sys.argv[1:])                                        # optional: regenerate the named cases only
    pick = lambda cases: {k: v for k, v in cases.items() if not only or k in only}
    global KMEANS_CASES, MI_CASES, BATCH_MI_CASES, AMI_CASES
    KMEANS_CASES, MI_CASES, BATCH_MI_CASES = pick(KMEANS_CASES), pick(MI_CASES), pick(BATCH_MI_CASES)
    AMI_CASES = pick(AMI_CASES)
    if not only or "ref_ver1_checkpoint" in only:
        write_reference_checkpoint(os.path.join(GOLDEN, "ref_ver1_cache_epoch_0.pkl"))
        print("ref_ver1_cache_epoch_0.pkl written by the reference's KMeans class")
    for name, case in KMEANS_CASES.items():
        out = run_reference_kmeans(case)
        np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
        print(name, "count", out["count"], "fallback", out["fallback"],
              "last dist", out["step_mean_dist"][-1], "ids", np.bincount(out["assign_best"]).tolist())
    for name, case in MI_CASES.items():
        for measure in ("mem_mi",) + (("mi",) if name == "mi_dense_small" else ()):
            out = run_reference_mi(case, measure)
            np.savez_compressed(os.path.join(GOLDEN, f"{name}_{measure}.npz"), **out)
            print(name, measure, "|S|", len(out["S"]), out["S"][:8].tolist(), out["GAIN"][:3].tolist())
    for name, case in AMI_CASES.items():
        out = run_reference_mi(case, "ami")
        np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
        print(name, "ami |S|", len(out["S"]), out["S"][:8].tolist(), out["GAIN"][:3].tolist())
    for name, case in BATCH_MI_CASES.items():
        out = run_reference_batch_mi(case)
        np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
        print(name, "batch_mi |S|", len(out["S"]), out["S"][:8].tolist(), out["GAIN"][:3].tolist())


if __name__ == "__main__":
    main()
