"""`cli run` with the reference's measure names over ALL ten layer clusterings (pairing 'combination', 45 contingency
tables): `mem_mi` against the oracle driver index for index, `ami` for the reference's output contract."""
import csv

import numpy as np
import pytest
import torch

from acav100m_b200.clustering import cli as ccli
from acav100m_b200.subset_selection import cli as scli, dataloader as sdata
from oracle import mi_oracle as mo
from tests.shard_fixtures import write_feature_shards

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def clustered(tmp_path_factory):
    root = tmp_path_factory.mktemp("measures")
    feat_dir, meta_dir = write_feature_shards(root / "data", n_shards=2, clips_per_shard=70, seed=8)
    clusters = root / "data" / "clusters"
    torch.manual_seed(4)
    ccli.main(["cluster", "--feature_path=" + str(feat_dir / "shard-{000000..000001}.pkl"),
               "--out_path=" + str(clusters), "--meta_path=" + str(meta_dir), "--clustering.ncentroids=6",
               "--data.batch_size=32", "--computation.num_gpus=1"])
    return root, clusters, meta_dir


def test_cli_mem_mi_over_all_pairs_matches_oracle_driver(clustered):
    root, clusters, meta_dir = clustered
    out_csv = root / "data" / "mem_mi.csv"
    glob = str(clusters / "shard-{000000..000001}.pkl")
    scli.main(["run", "--shards_path=" + glob, "--meta_path=" + str(meta_dir), "--out_path=" + str(out_csv),
               "--measure_name=mem_mi", "--subset.size=30", "--shuffle_candidates=False", "--verbose=False"])
    parts, _ = sdata.load_data(glob, meta_dir)
    a, shard_names, filenames, types = sdata.preprocess(parts[sorted(parts)[0]])
    assert a.shape == (140, 10)
    S, _ = mo.run_greedy_driver(a, subset_size=30, clustering_types=types, shuffle_candidates=False)
    lines = list(csv.reader(open(out_csv)))
    assert [l[1] for l in lines] == [filenames[s] for s in sorted(S)] and len(lines) == 29


def test_cli_ami_runs_on_all_pairs(clustered):
    root, clusters, meta_dir = clustered
    out_csv = root / "data" / "ami.csv"
    scli.main(["run", "--shards_path=" + str(clusters / "shard-{000000..000001}.pkl"), "--meta_path=" + str(meta_dir),
               "--out_path=" + str(out_csv), "--measure_name=ami", "--subset.size=25", "--shuffle_candidates=False",
               "--verbose=False"])
    lines = list(csv.reader(open(out_csv)))
    names = [l[1] for l in lines]
    assert len(lines) == 24 and len(set(names)) == 24 and all(n.endswith(".mp4") for n in names)   # mi.py:161: size - 1
    assert np.all([l[2].startswith("yt") for l in lines])


def test_run_sh_runs_both_stages_with_the_reference_defaults(tmp_path):
    """`bash run.sh DATA_DIR`: the reference's two stage scripts (clustering/code/run.sh, subset_selection/code/run.sh)
    with their default flags -- K = 32, two epochs at batch 32, then `batch_mi` over all 45 pairs -- on shard-000000."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    write_feature_shards(tmp_path / "data", n_shards=1, clips_per_shard=80, seed=12)
    env = dict(os.environ, PATH=os.path.dirname(sys.executable) + os.pathsep + os.environ.get("PATH", ""))
    out = subprocess.run(["bash", os.path.join(root, "run.sh"), str(tmp_path / "data"), "--subset.size=12",
                          "--verbose=False"], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert out.stdout.count("done") >= 2
    assert (tmp_path / "data" / "clusters" / "shard-000000.pkl").is_file()
    lines = list(csv.reader(open(tmp_path / "data" / "output.csv")))
    assert len(lines) == 12 and len({l[1] for l in lines}) == 12 and all(l[2].startswith("yt") for l in lines)
