"""Host-side I/O of the two CLI shims: reference on-disk formats in, reference formats out (CPU only)."""
import csv
import json
import os
import pickle
import sys
import types
from pathlib import Path

import numpy as np
import pytest
import torch

from acav100m_b200 import hostio
from acav100m_b200.clustering import args as cargs, data as cdata, save as csave
from acav100m_b200.clustering.config import MODELS, defaults as cdefaults
from acav100m_b200.subset_selection import cli as scli, dataloader as sdata, save as ssave
from acav100m_b200.subset_selection.config import defaults as sdefaults
from oracle import ref_shims
from tests.shard_fixtures import write_feature_shards


def test_cli_parsing_and_dotted_overrides():
    cmd, kw = hostio.parse_cli(["cluster", "--feature_path=d/shard-{000000..000002}.pkl", "--out_path=o",
                                "--clustering.ncentroids=8", "--data.batch_size", "64", "--debug"])
    assert cmd == "cluster" and kw["clustering.ncentroids"] == 8 and kw["data.batch_size"] == 64 and kw["debug"] is True
    args = cargs.get_args(**cargs.cli_aliases(kw))
    assert args.clustering.ncentroids == 8 and args.data.batch_size == 64
    assert str(args.data.path).endswith("d/shard-{000000..000002}.pkl")
    assert args.data.output.path == Path("o").resolve()
    assert args.clustering.save_scheme_ver2 is None and args.node_rank is None      # DefaultMunch(None) semantics
    assert hostio.braceexpand("a/shard-{000008..000011}.pkl") == ["a/shard-0000%02d.pkl" % i for i in (8, 9, 10, 11)]


@pytest.mark.skipif(not ref_shims.reference_available(), reason="/root/reference not mounted")
def test_default_flag_trees_equal_the_reference():
    for sub, ours in (("clustering", cdefaults), ("subset_selection", sdefaults)):
        spec = {}
        exec(open(os.path.join(ref_shims.REFERENCE_ROOT, sub, "code", "config.py")).read(), spec)
        ref = spec["defaults"]
        mine = {k: v for k, v in ours.items()}
        if sub == "subset_selection":
            mine = json.loads(json.dumps(mine))
            mine["clustering"].pop("columns")                  # our one documented extension
            ref = json.loads(json.dumps(ref))
        assert mine == ref


def test_feature_shards_collate_like_the_reference(tmp_path):
    feat_dir, _ = write_feature_shards(tmp_path, n_shards=2, clips_per_shard=10)
    paths = cdata.expand_shards(feat_dir / "shard-{000000..000001}.pkl")
    assert len(paths) == 2
    got = list(cdata.batches(paths, 8, drop_last=True))
    assert len(got) == 2                                        # 20 rows -> 2 full batches, batches span shards
    b = got[1]
    assert set(b.keys()) == {"VGGish/YouTube-8M", "SLOWFAST_8x8_R50/kinetics-400", "filename", "shard_name",
                             "shard_size", "idx"}
    assert b["shard_name"][:2] == ["shard-000000"] * 2 and b["shard_name"][2:] == ["shard-000001"] * 6
    for key, model in (("VGGish/YouTube-8M", "layer_vggish"), ("SLOWFAST_8x8_R50/kinetics-400", "layer_slow_fast")):
        for i, d in enumerate(MODELS[model]["output_dims"]):
            t = b[key]["layer_%d" % i]
            assert t.dtype == torch.float32 and tuple(t.shape) == (8, d)
    assert b["idx"][0] == Path(b["filename"][0]).stem
    half = cdata.rank_slice(b, 1, 2)
    assert half["idx"] == b["idx"][4:] and torch.equal(half["VGGish/YouTube-8M"]["layer_0"],
                                                       b["VGGish/YouTube-8M"]["layer_0"][4:])
    assert len(list(cdata.batches(paths, 8, drop_last=False))) == 3


def _fake_cluster_shard(args, name, n, rng):
    ids = ["clip_%s_%04d" % (name[-6:], c) for c in range(n)]
    data = []
    for m in ("layer_vggish", "layer_slow_fast"):
        per = {idx: {"assignments": {"layer_%d" % i: np.int64(rng.randint(8)) for i in range(5)},
                     "filename": idx + ".mp4", "shard_name": name, "shard_size": n} for idx in ids}
        data.append({"model_key": m, "data": per, **MODELS[m]["tag"]})
    return csave.save_assignments(args, name, ids, data)


def test_cluster_shards_round_trip_into_selection_inputs(tmp_path):
    rng = np.random.RandomState(0)
    args = cargs.get_args(**{"data.output.path": str(tmp_path / "clusters")})
    p0 = _fake_cluster_shard(args, "shard-000000", 6, rng)
    p1 = _fake_cluster_shard(args, "shard-000001", 5, rng)
    log = csave.store_shards_set(args, [p0, p1])
    assert log.name.startswith("log_") and json.load(open(log))["shards"] == ["shard-000000", "shard-000001"]
    row = pickle.load(open(p0, "rb"))[0]
    assert set(row) == {"video_assignments", "audio_assignments", "filename", "shard_size", "shard_name"}
    feat = row["audio_assignments"][0]
    assert feat["model_key"] == "layer_vggish" and feat["extractor_name"] == "VGGish"
    assert isinstance(feat["array"]["layer_3"], np.int64)
    parts, metas = sdata.load_data(str(tmp_path / "clusters" / "shard-{000000..000001}.pkl"), tmp_path)
    assert list(parts.keys()) == [0] and len(parts[0]) == 11 and metas == {}
    a, shard_names, filenames, ctypes_ = sdata.preprocess(parts[0])
    assert a.shape == (11, 10) and a.dtype == np.int64
    assert ctypes_ == sorted(ctypes_) and ctypes_[0] == ("layer_slow_fast", "layer_0") and ctypes_[-1] == ("layer_vggish", "layer_4")
    a2, _, _, t2 = sdata.preprocess(parts[0], columns=[("layer_vggish", "layer_4"), ("layer_slow_fast", "layer_4")])
    assert a2.shape == (11, 2) and np.array_equal(a2[:, 0], a[:, 9]) and np.array_equal(a2[:, 1], a[:, 4])
    if ref_shims.reference_available():                        # the reference's own reader agrees
        sys.modules.setdefault("braceexpand", types.SimpleNamespace(braceexpand=hostio.braceexpand))
        code = os.path.join(ref_shims.REFERENCE_ROOT, "subset_selection", "code")
        sys.path.insert(0, code)
        try:
            for stale in ("dataloader", "utils", "multiprocess"):
                sys.modules.pop(stale, None)
            import dataloader as ref_loader
            ra, rs, rf, rt = ref_loader.preprocess(parts[0], num_workers=1)
            assert np.array_equal(ra, a) and list(rs) == list(shard_names) and list(rf) == list(filenames) and rt == ctypes_
        finally:
            sys.path.remove(code)
            for stale in ("dataloader", "utils", "multiprocess"):
                sys.modules.pop(stale, None)


def test_stack_layer_equals_np_stack():
    rng = np.random.RandomState(3)
    dicts = [{"layer_0": rng.standard_normal(5).astype(np.float32), "layer_1": rng.standard_normal(3)} for _ in range(7)]
    lists = [[d["layer_0"], d["layer_1"]] for d in dicts]
    for arrays in (dicts, lists, dicts[:3] + lists[3:]):
        for layer in ("layer_0", "layer_1"):
            want = np.stack([np.asarray(cdata._get_layer(a, layer), dtype=np.float32) for a in arrays])
            got = cdata._stack_layer(arrays, layer)
            assert got.dtype == np.float32 and np.array_equal(got, want)


def test_file_stem_equals_pathlib():
    import random
    from pathlib import Path as P
    rnd = random.Random(0)
    names = ["a.mp4", "x/y/a.b.mp4", ".hidden", "name.", "noext", "a/b/", "", "..", "x/y.z/.", "C:\\a\\b.mp4", "a.tar.gz"]
    names += ["".join(rnd.choice("ab./_-") for _ in range(rnd.randint(0, 9))) for _ in range(3000)]
    assert [hostio.file_stem(n) for n in names] == [P(n).stem for n in names]


def test_preprocess_column_plan_equals_per_row_dicts():
    """`preprocess` reads rows through a column plan taken from the first row; rows laid out differently (feature order
    swapped, list- or scalar-valued 'array', dataloader.py:17-36) fall back to the per-row dict and give the same ids."""
    rng = np.random.RandomState(2)

    def row(i, swap=False, kind="dict"):
        def feat(model, n):
            vals = [np.int64(rng.randint(9)) for _ in range(n)]
            arr = {"layer_%d" % j: v for j, v in enumerate(vals)} if kind == "dict" else (vals if kind == "list" else vals[0])
            return {"model_key": model, "array": arr}
        video = [feat("layer_slow_fast", 3), feat("layer_other", 3)]
        return {"filename": "c%03d.mp4" % i, "shard_name": "shard-%06d" % (i // 4),
                "audio_assignments": [feat("layer_vggish", 3)], "video_assignments": video[::-1] if swap else video}

    for kind in ("dict", "list", "scalar"):
        data = [row(i, swap=(i % 3 == 1), kind=kind) for i in range(12)]
        a, shard_names, filenames, types_ = sdata.preprocess(data)
        want_rows = [sdata.format_row(r)[2] for r in data]
        assert types_ == sorted(want_rows[0].keys()) and len(types_) == (3 if kind == "scalar" else 9)
        assert np.array_equal(a, np.array([[int(r[k]) for k in types_] for r in want_rows]))
        assert a.dtype == np.int64 and filenames[5] == "c005.mp4" and shard_names[5] == "shard-000001"
        cols = [types_[-1], types_[0]]
        a2, _, _, t2 = sdata.preprocess(data, columns=cols)
        assert t2 == cols and np.array_equal(a2, a[:, [len(types_) - 1, 0]])
    with pytest.raises(AssertionError):
        sdata.preprocess(data, columns=[("layer_vggish", "nope")])


def test_compare_measures_command_plumbing(tmp_path, monkeypatch, capsys):
    """`cli compare_measures` (reference cli.py:80-83, tests.py): measure list from --measure_names, one report per
    partition; the oracle stands in for the device measures here (the real ones run in the GPU tests)."""
    from acav100m_b200.subset_selection import cli as scli, compare
    from oracle import mi_oracle as mo
    rng = np.random.RandomState(1)
    args = cargs.get_args(**{"data.output.path": str(tmp_path / "clusters")})
    paths = [_fake_cluster_shard(args, "shard-00000%d" % i, 9, rng) for i in range(2)]
    csave.store_shards_set(args, paths)
    calls = []

    def fake_run_greedy(a, assignments, clustering_types, subset_size, subset_ratio, measure_name, cluster_pairing,
                        shuffle_candidates, verbose):
        calls.append((measure_name, cluster_pairing, assignments.shape))
        S, GAIN = mo.run_greedy_driver(assignments, subset_size=subset_size, pairing=cluster_pairing,
                                       clustering_types=clustering_types)
        if measure_name == "mi":                                 # a second measure that disagrees on the last pick
            S, GAIN = S[:-1] + [S[0]], GAIN[:-1] + [GAIN[-1] + 0.5]
        return S, GAIN, [0.0] * len(GAIN)

    monkeypatch.setattr(compare, "_run_greedy", fake_run_greedy)
    report = scli.compare_measures(shards_path=str(tmp_path / "clusters" / "shard-{000000..000001}.pkl"),
                                   meta_path=str(tmp_path), out_path=str(tmp_path / "out.csv"),
                                   **{"subset.size": 6, "shuffle_candidates": False, "verbose": False,
                                      "clustering.columns": [("layer_vggish", "layer_4"), ("layer_slow_fast", "layer_4")]})
    assert [c[0] for c in calls] == ["mem_mi", "mi"] and calls[0][1] == "combination" and calls[0][2] == (18, 2)
    assert len(report) == 1 and report[0][0][:2] == ("mem_mi", "mi")
    assert report[0][0][2] == pytest.approx(4 / 5) and report[0][0][3] == pytest.approx(0.5 / 4)
    out = capsys.readouterr().out
    assert "mem_mi vs. mi" in out and "S equivalence" in out and out.strip().endswith("done")


def test_partitions_follow_newest_log(tmp_path):
    d = tmp_path / "c"
    d.mkdir()
    json.dump({"shards": ["shard-000000", "shard-000001"]}, open(d / "log_host_1_100.json", "w"))
    json.dump({"shards": ["shard-000001"]}, open(d / "log_host_2_200.json", "w"))
    assert sdata.load_partitions(d) == {"shard-000000": 0, "shard-000001": 1}


def test_output_csv_matches_reference_example(tmp_path):
    metas = {"shard-000000": {"a": {"filename": "a.mp4", "id": "qZ1", "segment": [0.0, 10.0]}}}
    data = [{"filename": "a.mp4", "shard_name": "shard-000000"}, {"filename": "b.mp4", "shard_name": "shard-000000"}]
    out, n = ssave.save_output(data, metas, tmp_path / "output.csv")
    out, n2 = ssave.save_output(data[:1], metas, tmp_path / "output.csv")            # append mode
    rows = list(csv.reader(open(out)))
    assert n == 2 and n2 == 1 and len(rows) == 3
    assert rows[0] == ["shard-000000", "a.mp4", "qZ1", "[0.0, 10.0]"]
    assert rows[1] == ["shard-000000", "b.mp4", "-1", "[-1.0, -1.0]"]
    if ref_shims.reference_available():
        ex = list(csv.reader(open(os.path.join(ref_shims.REFERENCE_ROOT, "examples", "output.csv"))))
        assert len(ex[0]) == 4 and ex[0][3].startswith("[") and ex[0][0].startswith("shard-")


def test_selection_cli_path_handling(tmp_path):
    (tmp_path / "clusters").mkdir()
    args = scli.prepare(out_path=str(tmp_path / "out"), shards_path=str(tmp_path / "clusters" / "shard-{0..1}.pkl"),
                        **{"measure_name": "mem_mi", "subset.ratio": 0.1})
    assert args.data.output.path == tmp_path / "out" / "output.csv"      # bare dir -> /output.csv
    assert args.data.meta.path == tmp_path / "clusters"                  # meta defaults to the shard dir
    assert args.measure_name == "mem_mi" and args.subset.ratio == 0.1 and args.computation.device == "cuda"
    assert args.batch.batch_size == 20 and args.log_times == 10


# ---- chunked selection (reference subset_selection/code/chunk.py) ---------------------------------------

def _chunk_fixture(tmp_path, n_shards=5, clips=7):
    rng = np.random.RandomState(1)
    cargs_ = cargs.get_args(**{"data.output.path": str(tmp_path / "clusters")})
    for s in range(n_shards):
        name = "shard-%06d" % s
        _fake_cluster_shard(cargs_, name, clips, rng)
        json.dump([{"filename": "clip_%06d_%04d.mp4" % (s, c), "id": "yt%d_%d" % (s, c), "segment": [0.0, 1.0]}
                   for c in range(clips)], open(tmp_path / "clusters" / (name + ".json"), "w"))
    return tmp_path / "clusters"


def test_chunk_planning_follows_the_reference(tmp_path):
    from acav100m_b200.subset_selection import chunk as schunk
    clusters = _chunk_fixture(tmp_path)
    assert list(schunk.get_chunks(list(range(5)), 2)) == [[0, 1], [2, 3], [4]]
    assert list(schunk.split_chunks(list(range(5)), 2)) == [[0, 1, 2], [3, 4]]
    args = scli.prepare(out_path=str(tmp_path / "out"), shards_path=str(clusters / "shard-{000000..000007}.pkl"),
                        **{"chunk_size": 2, "subset.size": 10, "computation.num_workers": 8})
    args.computation.num_gpus = 2
    chunk_args, nodes, num_chunks = schunk.plan_chunks(args)
    assert num_chunks == 3                                       # 5 existing shards of the 8 globbed, 2 per chunk
    assert chunk_args.subset.size == 4                           # ceil(10 / 3), chunk.py:45-46
    assert chunk_args.computation.num_workers == 4               # chunk.py:47-48
    assert [[num for num, _ in run] for run in nodes] == [[0, 1], [2]]
    assert [Path(p).name for p in nodes[1][0][1]] == ["shard-000004.pkl"]
    args.computation.num_gpus = 9                                # more GPUs than chunks -> thresholded, chunk.py:31-35
    _, nodes, _ = schunk.plan_chunks(args)
    assert args.computation.num_gpus == 3 and len(nodes) == 3


def test_chunk_caches_reduce_to_the_output_csv(tmp_path):
    from acav100m_b200.subset_selection import chunk as schunk
    clusters = _chunk_fixture(tmp_path)
    out = tmp_path / "out" / "output.csv"
    args = scli.prepare(out_path=str(out), shards_path=str(clusters / "shard-{000000..000004}.pkl"),
                        meta_path=str(clusters), **{"chunk_size": 2, "subset.size": 3, "verbose": False})
    args.parent_pid = "4242"
    parts, metas = sdata.load_data(str(clusters / "shard-{000000..000001}.pkl"), clusters)
    res = [{"filename": r["filename"], "shard_name": r["shard_name"]} for r in parts[sorted(parts)[0]][:5]]
    # pickled cache of chunk 0 and CSV cache of chunk 1 (the two --save_cache_as_csvs settings)
    schunk.save_chunk_cache(args, 0, 0, res, metas)
    cache = tmp_path / "out" / "caches" / "cache_4242_0_0.pkl"
    assert cache.is_file() and set(pickle.load(open(cache, "rb"))) == {"res", "metas"}
    p, n = schunk._reduce_single_cache(args, "cache_4242_0_1", res[::-1], metas)
    assert p.name == "cache_4242_0_1_output.csv" and n == 3      # truncated to the per-chunk quota
    rows = list(csv.reader(open(p)))
    assert rows[0][1] == res[-1]["filename"] and rows[0][2].startswith("yt")     # metadata joined by clip stem
    assert ssave.group_cache_paths([cache, p]) == {"cache_4242": [cache, p]}
    ssave.merge_all_csvs(args)                                   # `cli reduce_csvs`
    assert [r[1] for r in csv.reader(open(out))] == [r["filename"] for r in res[::-1][:3]]
    out.unlink()
    p.unlink()
    schunk.reduce_all_pkls(args)                                 # `cli reduce_pkls`
    assert [r[1] for r in csv.reader(open(out))] == [r["filename"] for r in res[:3]]
    assert (tmp_path / "out" / "caches" / "cache_4242_0_0_output.csv").is_file()


# ---- centroid checkpoints: both reference layouts, shard-subset lookup (run_clustering.py:55-116) ---------------

def test_reference_written_ver1_checkpoint_loads_without_the_reference(golden_dir):
    """tests/golden/ref_ver1_cache_epoch_0.pkl was written by the UNMODIFIED reference class (oracle/gen_golden.py::
    write_reference_checkpoint): object pickles under the class path sgd_clustering.KMeans.  It must load here with no
    `sgd_clustering` module importable."""
    import sys
    from acav100m_b200.clustering import checkpoint
    assert "sgd_clustering" not in sys.modules
    path = os.path.join(golden_dir, "ref_ver1_cache_epoch_0.pkl")
    tree = checkpoint.load_tree(path)
    want = dict(np.load(path + ".expect.npz"))
    assert sorted(tree) == ["layer_slow_fast", "layer_vggish"] and sorted(tree["layer_vggish"]) == ["layer_0", "layer_1"]
    for m, per in tree.items():
        for layer, attrs in per.items():
            key = "%s/%s/" % (m, layer)
            assert isinstance(attrs["centers"], np.ndarray) and attrs["centers"].dtype == np.float32
            assert np.array_equal(attrs["centers"], want[key + "centers"])
            assert np.array_equal(attrs["counts"], want[key + "counts"])
            assert attrs["count"] == int(want[key + "count"]) and attrs["fallback"] == int(want[key + "fallback"])
            assert attrs["initial_rounds"] == 10 and tuple(attrs["reinit"]) == (.7, 5.0) and attrs["sequential"] is False
    assert "sgd_clustering" not in sys.modules


def test_ver1_checkpoint_written_here_is_what_the_reference_class_unpickles(tmp_path):
    """save_scheme_ver2=False: objects pickled under `sgd_clustering.KMeans` carrying exactly the reference's attributes
    as CPU tensors -- what the reference's load path needs (torch.load, then .to(device) on every object,
    run_clustering.py:93, sgd_clustering.py:59-61).  Checked with a stand-in module of that name, and -- when
    /root/reference is mounted -- with the reference class itself."""
    import sys
    import types
    import torch
    from acav100m_b200.clustering import checkpoint
    from oracle import ref_shims
    tree = {"layer_vggish": {"layer_0": {"args": None, "count": 320, "lr": 0.01, "initial_rounds": 10, "reinit": (.7, 5.0),
                                         "fallback": 2, "sequential": False,
                                         "centers": np.arange(12, dtype=np.float32).reshape(3, 4),
                                         "counts": np.array([5, 0, 7], dtype=np.float32)}}}
    path = tmp_path / "cache_epoch_0_shard-{000000..000001}.pkl"
    checkpoint.save_tree_ver1(tree, path)
    assert "sgd_clustering" not in sys.modules
    # round trip through our own reader
    back = checkpoint.load_tree(path)["layer_vggish"]["layer_0"]
    assert np.array_equal(back["centers"], tree["layer_vggish"]["layer_0"]["centers"]) and back["count"] == 320

    def check(cls):
        objs = torch.load(str(path), weights_only=False)
        km = objs["layer_vggish"]["layer_0"]
        assert type(km) is cls
        assert torch.is_tensor(km.centers) and km.centers.dtype == torch.float32 and km.centers.shape == (3, 4)
        assert torch.equal(km.counts, torch.tensor([5., 0., 7.])) and km.count == 320 and km.fallback == 2
        assert km.lr == 0.01 and km.initial_rounds == 10 and tuple(km.reinit) == (.7, 5.0) and km.sequential is False
        return km

    mod = types.ModuleType("sgd_clustering")
    mod.KMeans = type("KMeans", (), {"__module__": "sgd_clustering"})
    sys.modules["sgd_clustering"] = mod
    try:
        check(mod.KMeans)
    finally:
        del sys.modules["sgd_clustering"]
    if ref_shims.reference_available():
        RefKMeans = ref_shims.load_reference_kmeans()
        try:
            km = check(RefKMeans)
            km.to("cpu")                                           # sgd_clustering.py:59-61
            best, _ = km.calc_best(torch.zeros(2, 4))             # past warm-up (count >= 10 * k): distance branch
            assert best.tolist() == [0, 0]
            assert set(km.get_attrs()) >= {"centers", "counts", "count", "lr", "fallback"}
        finally:
            sys.modules.pop("sgd_clustering", None)


def test_shard_subset_cache_lookup(tmp_path):
    """get_shard_subset_cache (run_clustering.py:76-84): a checkpoint trained on a subset of the requested shards."""
    import types
    from acav100m_b200.clustering import run_clustering as rc
    for name in ("cache_epoch_1_shard-{000000..000001}.pkl", "cache_epoch_1_shard-{000000..000003}.pkl",
                 "cache_epoch_1_shard-{000004..000009}.pkl", "cache_epoch_2_shard-{000000..000001}.pkl", "log_x.json"):
        (tmp_path / name).write_bytes(b"")
    args = types.SimpleNamespace()
    got = rc.get_shard_subset_cache(args, tmp_path, 1, "shard-{000000..000003}.pkl")
    assert got is not None and got.name in ("cache_epoch_1_shard-{000000..000001}.pkl", "cache_epoch_1_shard-{000000..000003}.pkl")
    # only subsets qualify: shards 4..9 are not inside 0..3; epoch must match
    assert rc.get_shard_subset_cache(args, tmp_path, 1, "shard-{000002..000003}.pkl") is None
    assert rc.get_shard_subset_cache(args, tmp_path, 3, "shard-{000000..000009}.pkl") is None
    got2 = rc.get_shard_subset_cache(args, tmp_path, 2, "shard-{000000..000009}.pkl")
    assert got2.name == "cache_epoch_2_shard-{000000..000001}.pkl"


# ---- parallel shard loader (clustering/loader.py): same batches as the in-process reader ------------------------------

@pytest.mark.parametrize("workers,batch_size,drop_last", [(0, 7, False), (2, 7, False), (3, 64, True), (2, 100, False),
                                                          (2, 23, True)])
def test_parallel_loader_yields_the_in_process_batch_sequence(tmp_path, workers, batch_size, drop_last):
    """Worker processes unpickle + collate whole shards into shared memory; the parent must cut exactly the batches of
    data.batches (shard order, row order, batches straddling shards, ragged tail, drop_last)."""
    import torch
    from acav100m_b200.clustering import data as cdata, loader
    from tests.shard_fixtures import write_feature_shards
    feat_dir, _ = write_feature_shards(tmp_path / "d", n_shards=5, clips_per_shard=30, seed=2)
    paths = cdata.expand_shards(str(feat_dir / "shard-{000000..000004}.pkl"))
    want = list(cdata.batches(paths, batch_size, drop_last))
    got = list(loader.ShardLoader(paths, batch_size, drop_last, workers=workers, hold=10))
    assert len(got) == len(want) > 0
    for g, w in zip(got, want):
        assert g.keys() == w.keys()
        for k, v in w.items():
            if isinstance(v, dict):
                for layer, t in v.items():
                    assert torch.equal(g[k][layer], t), (k, layer)
            elif torch.is_tensor(v):
                assert torch.equal(g[k], v)
            else:
                assert g[k] == v, k


def test_parallel_loader_skips_unreadable_shards_and_cleans_up(tmp_path):
    from acav100m_b200.clustering import data as cdata, loader
    from tests.shard_fixtures import write_feature_shards
    feat_dir, _ = write_feature_shards(tmp_path / "d", n_shards=3, clips_per_shard=10, seed=5)
    (feat_dir / "shard-000001.pkl").write_bytes(b"not a pickle")
    paths = cdata.expand_shards(str(feat_dir / "shard-{000000..000002}.pkl"))
    before = set(os.listdir("/dev/shm")) if os.path.isdir("/dev/shm") else set()
    n = sum(len(b["idx"]) for b in loader.ShardLoader(paths, 8, False, workers=2))
    assert n == 20
    if os.path.isdir("/dev/shm"):
        assert set(os.listdir("/dev/shm")) <= before, "shared-memory blocks must be unlinked"
