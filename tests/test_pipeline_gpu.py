"""End-to-end drop-in check on a GPU: feature shards -> `cli cluster` -> cluster shards -> `cli run`
-> output.csv, compared with an oracle replay of the same pipeline (same seeds, same batch order)."""
import csv
import pickle
import random

import numpy as np
import pytest
import torch

from acav100m_b200.clustering import cli as ccli, data as cdata
from acav100m_b200.clustering.config import MODELS
from acav100m_b200.subset_selection import cli as scli, dataloader as sdata
from oracle import kmeans_oracle as ko, mi_oracle as mo
from tests.shard_fixtures import write_feature_shards

pytestmark = pytest.mark.gpu


def test_cli_pipeline_end_to_end(tmp_path):
    feat_dir, meta_dir = write_feature_shards(tmp_path / "data", n_shards=3, clips_per_shard=100, seed=3)
    clusters, out_csv = tmp_path / "data" / "clusters", tmp_path / "data" / "output.csv"
    glob = str(feat_dir / "shard-{000000..000002}.pkl")
    k, bs, epochs = 8, 64, 2

    torch.manual_seed(11)
    ccli.main(["cluster", "--feature_path=" + glob, "--out_path=" + str(clusters), "--meta_path=" + str(meta_dir),
               "--clustering.ncentroids=%d" % k, "--data.batch_size=%d" % bs, "--clustering.epochs=%d" % epochs,
               "--computation.num_gpus=1"])
    shard_files = sorted(clusters.glob("shard-*.pkl"))
    assert [p.name for p in shard_files] == ["shard-00000%d.pkl" % i for i in range(3)]
    assert len(list(clusters.glob("log_*.json"))) == 1
    assert (clusters / "cache_epoch_1_shard-{000000..000002}.pkl").is_file()

    # oracle replay: same construction order, same global RNG stream, same batches
    torch.manual_seed(11)
    names = ["layer_vggish", "layer_slow_fast"]
    states = {m: {"layer_%d" % i: ko.new_state(d, k) for i, d in enumerate(MODELS[m]["output_dims"])} for m in names}
    keymap = {m: "/".join((MODELS[m]["tag"]["name"], MODELS[m]["tag"]["dataset"])) for m in names}
    paths = cdata.expand_shards(glob)
    for epoch in range(epochs):
        for batch in cdata.batches(paths, bs, drop_last=True):
            for m in names:
                for layer, st in states[m].items():
                    st.lr = ko.epoch_lr(epoch)
                    ko.sgd_step(st, batch[keymap[m]][layer])
    ckpt = torch.load(str(clusters / "cache_epoch_1_shard-{000000..000002}.pkl"), weights_only=False)
    for m in names:
        for layer, st in states[m].items():
            got = ckpt[m][layer]
            assert got["count"] == st.count
            np.testing.assert_allclose(got["centers"], st.centers.numpy(), rtol=2e-5, atol=1e-7)
            assert np.array_equal(got["counts"], st.counts.numpy())

    # cluster shards hold the oracle's assignments (up to fp32 near-ties)
    rows = [r for p in shard_files for r in pickle.load(open(p, "rb"))]
    assert len(rows) == 300 and rows[0]["filename"] == "clip_000000_0000.mp4"
    all_batch = next(cdata.batches(paths, 300, drop_last=False))
    agree, total = 0, 0
    for m in names:
        for layer, st in states[m].items():
            want, _ = ko.assign(st, all_batch[keymap[m]][layer])
            field = "audio_assignments" if "vggish" in m else "video_assignments"
            got = np.array([int(r[field][0]["array"][layer]) for r in rows])
            agree += int((got == want.numpy()).sum())
            total += len(got)
    assert agree / total > 0.995

    # selection through the CLI on one audio-visual pair, shuffle off
    cols = "[('layer_vggish','layer_4'),('layer_slow_fast','layer_4')]"
    scli.main(["run", "--shards_path=" + str(clusters / "shard-{000000..000002}.pkl"), "--meta_path=" + str(meta_dir),
               "--out_path=" + str(out_csv), "--measure_name=mem_mi", "--subset.ratio=0.2",
               "--shuffle_candidates=False", "--clustering.columns=" + cols, "--verbose=False"])
    lines = list(csv.reader(open(out_csv)))
    parts, _ = sdata.load_data(str(clusters / "shard-{000000..000002}.pkl"), meta_dir)
    a, shard_names, filenames, _ = sdata.preprocess(parts[0], columns=[("layer_vggish", "layer_4"),
                                                                       ("layer_slow_fast", "layer_4")])
    S, _ = mo.run_greedy_driver(a, subset_ratio=0.2, shuffle_candidates=False)
    want = sorted(S)
    assert len(lines) == len(want) == round(0.2 * 300) - 1
    assert [l[1] for l in lines] == [filenames[s] for s in want]
    assert lines[0][0] == shard_names[want[0]] and lines[0][2].startswith("yt") and lines[0][3].startswith("[")

    # shuffled candidates follow python's `random` like the reference (run_greedy.py:37-40)
    out2 = tmp_path / "data" / "out2.csv"
    random.seed(5)
    scli.main(["run", "--shards_path=" + str(clusters / "shard-{000000..000002}.pkl"), "--meta_path=" + str(meta_dir),
               "--out_path=" + str(out2), "--measure_name=mem_mi", "--subset.size=40", "--clustering.columns=" + cols,
               "--verbose=False"])
    S2, _ = mo.run_greedy_driver(a, subset_size=40, shuffle_candidates=True, rng=random.Random(5))
    assert [l[1] for l in csv.reader(open(out2))] == [filenames[s] for s in sorted(S2)]


def test_cli_default_measure_runs_on_all_pairs(tmp_path):
    """`cli run` with the reference's defaults: measure batch_mi, pairing 'combination' over all ten layer
    clusterings (45 contingency tables), ratio 0.2, shuffled candidates."""
    feat_dir, meta_dir = write_feature_shards(tmp_path / "data", n_shards=2, clips_per_shard=60, seed=4)
    clusters, out_csv = tmp_path / "data" / "clusters", tmp_path / "data" / "output.csv"
    torch.manual_seed(3)
    ccli.main(["cluster", "--feature_path=" + str(feat_dir / "shard-{000000..000001}.pkl"),
               "--out_path=" + str(clusters), "--meta_path=" + str(meta_dir), "--clustering.ncentroids=6",
               "--data.batch_size=32", "--computation.num_gpus=1"])
    random.seed(1)
    scli.main(["run", "--shards_path=" + str(clusters / "shard-{000000..000001}.pkl"),
               "--meta_path=" + str(meta_dir), "--out_path=" + str(out_csv), "--verbose=False"])
    lines = list(csv.reader(open(out_csv)))
    assert len(lines) == round(0.2 * 120)
    names = [l[1] for l in lines]
    assert len(set(names)) == len(names) and all(n.endswith(".mp4") for n in names)
    assert all(l[2].startswith("yt") for l in lines)


def test_cli_chunked_selection_and_reduce(tmp_path):
    """`cli run --chunk_size=2` (reference chunk.py): every chunk of two shards is its own greedy selection
    with quota ceil(size / num_chunks), cached as a CSV; `cli reduce` appends the caches to output.csv."""
    feat_dir, meta_dir = write_feature_shards(tmp_path / "data", n_shards=4, clips_per_shard=50, seed=6)
    clusters, out_csv = tmp_path / "data" / "clusters", tmp_path / "data" / "sel" / "output.csv"
    torch.manual_seed(2)
    ccli.main(["cluster", "--feature_path=" + str(feat_dir / "shard-{000000..000003}.pkl"),
               "--out_path=" + str(clusters), "--meta_path=" + str(meta_dir), "--clustering.ncentroids=6",
               "--data.batch_size=32", "--computation.num_gpus=1"])
    cols = [("layer_vggish", "layer_4"), ("layer_slow_fast", "layer_4")]
    common = ["--shards_path=" + str(clusters / "shard-{000000..000003}.pkl"), "--meta_path=" + str(meta_dir),
              "--out_path=" + str(out_csv), "--verbose=False"]
    scli.main(["run", *common, "--measure_name=mem_mi", "--subset.size=30", "--shuffle_candidates=False",
               "--chunk_size=2", "--computation.num_gpus=1", "--computation.load_async=True",
               "--clustering.columns=" + repr(cols).replace(" ", "")])
    caches = sorted((out_csv.parent / "caches").glob("cache_*_0_*_output.csv"))
    assert len(caches) == 2 and not out_csv.exists()
    want_all = []
    for i, cache in enumerate(caches):
        shard_glob = str(clusters / ("shard-{00000%d..00000%d}.pkl" % (2 * i, 2 * i + 1)))
        parts, _ = sdata.load_data(shard_glob, meta_dir)
        a, shard_names, filenames, _ = sdata.preprocess(parts[sorted(parts)[0]], columns=cols)
        S, _ = mo.run_greedy_driver(a, subset_size=15, shuffle_candidates=False)       # ceil(30 / 2) per chunk
        want = [filenames[s] for s in sorted(S)]
        got = [l[1] for l in csv.reader(open(cache))]
        assert got == want and len(got) == 14                   # run_greedy returns size - 1 picks (mi.py:150-192)
        want_all += want
    scli.main(["reduce", *common, "--subset.size=30"])
    lines = list(csv.reader(open(out_csv)))
    assert [l[1] for l in lines] == want_all and all(l[2].startswith("yt") for l in lines)
