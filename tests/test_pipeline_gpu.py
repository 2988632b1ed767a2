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


def _ckpt(path):
    return torch.load(str(path), weights_only=False)


def test_cli_resume_from_cached_epoch(tmp_path):
    """--clustering.cached_epoch with and without --clustering.resume_training (run_clustering.py:55-73,132-177,
    180-272): a loaded checkpoint skips training unless resume_training is set; resumed epochs continue the epoch
    counter (and with it the lr schedule); cluster shards are prefixed `epoch_{e}_`; shards already written are
    skipped; a checkpoint trained on a SUBSET of the shards is found through get_shard_subset_cache; a checkpoint in
    the reference's own (object) layout resumes like ours."""
    feat_dir, meta_dir = write_feature_shards(tmp_path / "data", n_shards=3, clips_per_shard=64, seed=9)
    glob = str(feat_dir / "shard-{000000..000002}.pkl")
    base = ["--meta_path=" + str(meta_dir), "--clustering.ncentroids=6", "--data.batch_size=32", "--computation.num_gpus=1"]
    name = "shard-{000000..000002}.pkl"

    # reference run: 3 epochs straight
    full = tmp_path / "full"
    torch.manual_seed(4)
    ccli.main(["cluster", "--feature_path=" + glob, "--out_path=" + str(full), *base, "--clustering.epochs=3"])
    want = _ckpt(full / ("cache_epoch_2_" + name))

    # 1 epoch, then resume for 2 more: same centers bit for bit (same data order, same lr per epoch)
    part = tmp_path / "part"
    torch.manual_seed(4)
    ccli.main(["cluster", "--feature_path=" + glob, "--out_path=" + str(part), *base, "--clustering.epochs=1"])
    assert (part / ("cache_epoch_0_" + name)).is_file() and (part / "shard-000000.pkl").is_file()
    first = _ckpt(part / ("cache_epoch_0_" + name))              # one epoch (the resumed run rewrites this file)
    ccli.main(["cluster", "--feature_path=" + glob, "--out_path=" + str(part), *base, "--clustering.epochs=2",
               "--clustering.cached_epoch=0", "--clustering.resume_training=True"])
    # the reference's loop is range(pre_epochs, epochs + pre_epochs) with pre_epochs = cached_epoch (:164): it re-runs
    # epoch 0's number, so two resumed epochs are numbered 0 and 1
    got = _ckpt(part / ("cache_epoch_1_" + name))
    for m in want:
        for layer in want[m]:
            assert got[m][layer]["count"] == want[m][layer]["count"]
            assert np.array_equal(got[m][layer]["centers"], want[m][layer]["centers"]), (m, layer)
    assert sorted(p.name for p in part.glob("epoch_0_shard-*.pkl")) == ["epoch_0_shard-00000%d.pkl" % i for i in range(3)]

    # cached_epoch without resume_training: no training (checkpoint untouched), assignment only, existing shards skipped
    stamp = (part / ("cache_epoch_1_" + name)).stat().st_mtime_ns
    (part / "epoch_1_shard-000001.pkl").write_bytes(b"sentinel")
    ccli.main(["cluster", "--feature_path=" + glob, "--out_path=" + str(part), *base, "--clustering.epochs=5",
               "--clustering.cached_epoch=1"])
    assert (part / ("cache_epoch_1_" + name)).stat().st_mtime_ns == stamp
    assert not (part / ("cache_epoch_2_" + name)).exists()
    assert (part / "epoch_1_shard-000001.pkl").read_bytes() == b"sentinel"          # :248-250 skip
    rows = pickle.load(open(part / "epoch_1_shard-000000.pkl", "rb"))
    rows_full = pickle.load(open(full / "shard-000000.pkl", "rb"))
    assert [r["filename"] for r in rows] == [r["filename"] for r in rows_full]
    same = sum(int(a["audio_assignments"][0]["array"]["layer_4"] == b["audio_assignments"][0]["array"]["layer_4"])
               for a, b in zip(rows, rows_full))
    assert same == len(rows)                                     # same centers -> same labels

    # a checkpoint trained on shards 0..1 serves a run over shards 0..2 (load_cache_from_shard_subset, default True)
    sub = tmp_path / "sub"
    torch.manual_seed(4)
    ccli.main(["cluster", "--feature_path=" + str(feat_dir / "shard-{000000..000001}.pkl"), "--out_path=" + str(sub),
               *base, "--clustering.epochs=1"])
    before = _ckpt(sub / "cache_epoch_0_shard-{000000..000001}.pkl")
    ccli.main(["cluster", "--feature_path=" + glob, "--out_path=" + str(sub), *base, "--clustering.cached_epoch=0"])
    assert not (sub / ("cache_epoch_0_" + name)).exists()        # loaded the subset checkpoint, trained nothing
    assert (sub / "epoch_0_shard-000002.pkl").is_file()
    after = _ckpt(sub / "cache_epoch_0_shard-{000000..000001}.pkl")
    assert np.array_equal(before["layer_vggish"]["layer_0"]["centers"], after["layer_vggish"]["layer_0"]["centers"])

    # the reference's default layout (objects under sgd_clustering.KMeans): written with save_scheme_ver2=False,
    # resumed from like ours
    v1 = tmp_path / "v1"
    torch.manual_seed(4)
    ccli.main(["cluster", "--feature_path=" + glob, "--out_path=" + str(v1), *base, "--clustering.epochs=1",
               "--clustering.save_scheme_ver2=False"])
    from acav100m_b200.clustering import checkpoint
    tree = checkpoint.load_tree(v1 / ("cache_epoch_0_" + name))
    assert np.array_equal(tree["layer_slow_fast"]["layer_2"]["centers"], first["layer_slow_fast"]["layer_2"]["centers"])
    ccli.main(["cluster", "--feature_path=" + glob, "--out_path=" + str(v1), *base, "--clustering.epochs=2",
               "--clustering.cached_epoch=0", "--clustering.resume_training=True", "--clustering.save_scheme_ver2=False"])
    got = checkpoint.load_tree(v1 / ("cache_epoch_1_" + name))
    for m in want:
        for layer in want[m]:
            assert np.array_equal(got[m][layer]["centers"], want[m][layer]["centers"]), (m, layer)
