"""Synthetic on-disk fixtures in the reference's formats."""
import json
import pickle
from pathlib import Path

import numpy as np

from acav100m_b200.clustering.config import MODELS


def write_feature_shards(root, n_shards=3, clips_per_shard=40, seed=0, models=('layer_vggish', 'layer_slow_fast')):
    """data/features/shard-NNNNNN.pkl + data/videos/shard-NNNNNN.json as feature_extraction writes them
    (feature_extraction/code/save.py:48-74)."""
    rng = np.random.RandomState(seed)
    feat_dir, meta_dir = Path(root) / "features", Path(root) / "videos"
    feat_dir.mkdir(parents=True, exist_ok=True)
    meta_dir.mkdir(parents=True, exist_ok=True)
    n_comp = 5
    protos = {m: [rng.standard_normal((n_comp, d)).astype(np.float32) * 3 for d in MODELS[m]['output_dims']]
              for m in models}
    for s in range(n_shards):
        name = "shard-%06d" % s
        rows, metas = [], []
        for c in range(clips_per_shard):
            fname = "clip_%06d_%04d.mp4" % (s, c)
            comp = rng.randint(n_comp)
            row = {'filename': fname, 'shard_name': name, 'shard_size': clips_per_shard,
                   'video_features': [], 'audio_features': []}
            for m in models:
                arr = {'layer_%d' % i: (protos[m][i][comp] + rng.standard_normal(d).astype(np.float32))
                       for i, d in enumerate(MODELS[m]['output_dims'])}
                feat = {'model_key': m, 'extractor_name': MODELS[m]['tag']['name'],
                        'dataset': MODELS[m]['tag']['dataset'], 'array': arr}
                row['audio_features' if 'vggish' in m else 'video_features'].append(feat)
            rows.append(row)
            metas.append({'filename': fname, 'id': "yt%05d" % (s * 1000 + c), 'segment': [float(c), float(c) + 10.0]})
        with open(feat_dir / (name + ".pkl"), "wb") as f:
            pickle.dump(rows, f)
        with open(meta_dir / (name + ".json"), "w") as f:
            json.dump(metas, f)
    return feat_dir, meta_dir
