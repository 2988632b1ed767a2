"""Host loops of the clustering stage (reference clustering/code/run_clustering.py:25-272 and
process_batch.py:6-69): train the per-(model, layer) ``KMeans`` objects over the feature shards, then
assign every clip and write cluster shards.  All tensor work happens inside ``KMeans`` (CUDA)."""
import copy
import math
from collections import defaultdict
from pathlib import Path

import torch

from .. import hostio
from . import checkpoint
from . import data as D
from .config import MODELS
from .save import save_assignments
from .sgd_clustering import KMeans


def filter_models(args):
    """utils.py:17-23."""
    names = list(args.models)
    data_name = Path(args.data.media.path).stem
    if data_name in args.data.types and args.data.types[data_name] == 'audio_only':
        names = [n for n in names if n in args.model_types.audio]
    return names


def model_key_map(model_names):
    """run_clustering.py:119-129 -- model name -> 'EXTRACTOR/dataset' key of the collated batch."""
    return {n: '/'.join((MODELS[n]['tag']['name'], MODELS[n]['tag']['dataset'])) for n in model_names}


def _world(args):
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def init_clusterings(args, model_names):
    """run_clustering.py:32-52."""
    clusterings = {}
    for name in model_names:
        dims = MODELS[name]['output_dims']
        if isinstance(dims, int):
            clusterings[name] = {'model': KMeans(args, dims, args.clustering.ncentroids)}
        else:
            clusterings[name] = {'layer_{}'.format(i): KMeans(args, d, args.clustering.ncentroids)
                                 for i, d in enumerate(dims)}
    return _place(args, clusterings)


def _place(args, clusterings):
    for per_model in clusterings.values():
        for km in per_model.values():
            km.to(args.computation.device)
            km.initialize()
    return clusterings


def cache_path(args, epoch):
    """utils.py:30-32 + run_clustering.py:110-116 -- ``cache_epoch_{e}_{basename(feature_path)}``."""
    return args.data.output.path / "cache_epoch_{}_{}".format(epoch, Path(args.data.path).name)


def save_clusterings(args, epoch, clusterings):
    """Checkpoint after every epoch (run_clustering.py:110-116).  ``--clustering.save_scheme_ver2`` (default True here):
    a tree of ``get_attrs()`` dicts with numpy arrays, which pickles no class; False writes the reference's own default
    layout (objects under the class path ``sgd_clustering.KMeans``) so that a reference run can resume from it."""
    tree = {m: {k: km.get_attrs() for k, km in per.items()} for m, per in clusterings.items()}
    for per in tree.values():
        for attrs in per.values():
            attrs['args'] = None
    path = cache_path(args, epoch)
    path.parent.mkdir(parents=True, exist_ok=True)
    if getattr(args.clustering, 'save_scheme_ver2', True) is False:
        checkpoint.save_tree_ver1(tree, path)
    else:
        torch.save(tree, str(path))
    return path


def get_shard_subset_cache(args, path, epoch, name):
    """run_clustering.py:76-84 -- a checkpoint trained on a SUBSET of the shards asked for now: of the files
    ``cache_epoch_{e}_*shard-{..}.pkl`` whose shard set lies inside ``name``'s, the reference takes the maximum by python's
    set comparison (``max(..., key=set)``: the first candidate, in glob order, that no later one is a proper superset of);
    kept as is.  Returns a path or None."""
    cands = {}
    for p in path.glob("cache_epoch_{}_*.pkl".format(epoch)):
        at = p.name.find('shard-')
        if at >= 0:
            cands[p] = set(hostio.braceexpand(p.name[at:]))
    shard_set = set(hostio.braceexpand(name))
    cands = {p: v for p, v in cands.items() if len(v - shard_set) == 0}
    if not cands:
        return None
    return max(list(cands.items()), key=lambda v: v[1])[0]


def _load_path(args, path, model_names):
    """run_clustering.py:87-107 -- ver1 (objects) and ver2 (dict trees) alike."""
    tree = checkpoint.load_tree(path)
    if not set(model_names) <= set(tree.keys()):
        print("clustering cache features does not match with the given models")
        return None
    clusterings = {}
    for m in model_names:
        clusterings[m] = {}
        for k, attrs in tree[m].items():
            km = KMeans.load(dict(attrs))
            km.args = args
            clusterings[m][k] = km
    print("loading from clustering cache: {}".format(path))
    return _place_loaded(args, clusterings)


def _place_loaded(args, clusterings):
    # the reference calls initialize() on loaded models too (run_clustering.py:49-52,95): with several ranks that
    # averages the (identical) replicas, a no-op up to rounding; kept
    return _place(args, clusterings)


def load_clusterings(args, model_names):
    """run_clustering.py:55-73 -- resume from ``--clustering.cached_epoch``."""
    epoch = args.clustering.cached_epoch
    if isinstance(epoch, int):
        path = cache_path(args, epoch)
        loaded = None
        if path.is_file():
            loaded = _load_path(args, path, model_names)
        elif getattr(args.clustering, 'load_cache_from_shard_subset', False):
            sub = get_shard_subset_cache(args, path.parent, epoch, Path(args.data.path).name)
            if sub is not None:
                loaded = _load_path(args, sub, model_names)
        if loaded is not None:
            return loaded, True
        print("no clustering cache found.")
    return init_clusterings(args, model_names), False


class _Stager:
    """Double-buffered device copies of the collated batches: one non-blocking H2D per (model, layer) on a copy stream
    into slot i % 2, so that batch i+1 is on its way while the KMeans objects (each on its own stream) work on batch i.
    The slots have fixed addresses, which is what lets ``KMeans.add`` replay its step from a CUDA graph."""

    def __init__(self, device):
        self.device = torch.device(device)
        self.copy = torch.cuda.Stream(self.device)
        self.slots, self.done, self.i = [{}, {}], [[], []], 0

    def _put(self, slot, key, t):
        buf = self.slots[slot].get(key)
        if buf is None or buf.shape != t.shape:
            buf = torch.empty(t.shape, dtype=torch.float32, device=self.device)
            self.slots[slot][key] = buf
        buf.copy_(t, non_blocking=True)
        return buf

    def upload(self, feats):
        """feats: {model: {layer: CPU tensor} | CPU tensor} -> (same tree on the device, ready event, slot)."""
        slot = self.i % 2
        self.i += 1
        with torch.cuda.stream(self.copy):
            for ev in self.done[slot]:                     # the steps that read this slot two batches ago
                self.copy.wait_event(ev)
            self.done[slot] = []
            out = {}
            for name, f in feats.items():
                if isinstance(f, dict):
                    out[name] = {layer: self._put(slot, (name, layer), t) for layer, t in f.items()}
                else:
                    out[name] = self._put(slot, (name, None), f)
            ready = torch.cuda.Event()
            ready.record(self.copy)
        return out, ready, slot

    def consumed(self, slot, stream):
        ev = torch.cuda.Event()
        ev.record(stream)
        self.done[slot].append(ev)


def _train_batch(args, features, clusterings, streams=None, ready=None, stager=None, slot=0, name=None):
    """process_batch.py:6-17 -- every (model, layer) KMeans takes its step; with `streams` each on its own CUDA stream."""
    # the reference collects the mean distances and drops them (:171-175); skipping them saves a pass
    # over the batch and the host sync of `.item()`
    items = clusterings.items() if isinstance(features, dict) else [('model', clusterings['model'])]
    for key, km in items:
        x = features[key] if isinstance(features, dict) else features
        if streams is None:
            km.add(x, sync=False, distance=False)
            continue
        s = streams.get((name, key))
        if s is None:
            s = streams[(name, key)] = torch.cuda.Stream(x.device)
        s.wait_event(ready)
        with torch.cuda.stream(s):
            km.add(x, sync=False, distance=False)
        stager.consumed(slot, s)
    return None


def _loader_workers(args, n_shards):
    """Worker processes of the shard loader: the reference's ``--computation.num_workers`` (default 40,
    data/clustering.py:128-136), capped by the shards and the cores of this host."""
    import os
    want = args.computation.num_workers
    want = 4 if want is None else int(want)
    return max(0, min(want, n_shards, (os.cpu_count() or 2) - 1))


def train_clusters(args, model_names):
    """run_clustering.py:132-177.  Shards are unpickled and collated by worker processes into page-locked shared memory
    (clustering/loader.py), copied to the device one batch ahead on a copy stream, and the per-(model, layer) KMeans
    objects step on their own streams."""
    from .loader import ShardLoader
    clusterings, loaded = load_clusterings(args, model_names)
    if loaded and not args.clustering.resume_training:
        return clusterings
    rank, world = _world(args)
    keys = model_key_map(model_names)
    shard_paths = D.expand_shards(args.data.path)
    pre_epochs = copy.deepcopy(args.clustering.cached_epoch) if loaded else 0
    epochs = math.ceil(args.clustering.epochs / max(args.computation.num_gpus or 1, 1))      # :146
    on_gpu = str(args.computation.device).startswith('cuda')
    stager = _Stager(torch.device('cuda', torch.cuda.current_device())) if on_gpu else None
    streams = {} if on_gpu else None
    print("training sgd kmeans for models: {}".format(model_names))
    for epoch in range(pre_epochs, epochs + pre_epochs):
        for per_model in clusterings.values():
            for km in per_model.values():
                km.lr = 0.1 ** (2 + epoch // 5)                                              # :168
        loader = ShardLoader(shard_paths, args.data.batch_size, drop_last=True,
                             workers=_loader_workers(args, len(shard_paths)), pin=on_gpu)
        for batch in loader:
            batch = D.rank_slice(batch, rank, world)
            feats = {name: batch[keys[name]] for name in model_names}
            if stager is None:
                for name in model_names:
                    _train_batch(args, feats[name], clusterings[name])
                continue
            dev, ready, slot = stager.upload(feats)
            for name in model_names:
                _train_batch(args, dev[name], clusterings[name], streams, ready, stager, slot, name)
        if on_gpu:
            torch.cuda.synchronize()
        if rank == 0:
            save_clusterings(args, epoch, clusterings)
    return clusterings


def _extract_batch(features, clusterings):
    """process_batch.py:37-56 -- ids per layer as np.int64."""
    if isinstance(features, dict):
        ids = {key: km.calc_best(features[key], sync=False, distance=False)[0] for key, km in clusterings.items()}
        ids = {key: v.cpu().numpy() for key, v in ids.items()}
        keys = sorted(ids.keys())
        return [dict(zip(keys, vals)) for vals in zip(*[ids[k] for k in keys])]
    return list(clusterings['model'].calc_best(features, sync=False, distance=False)[0].cpu().numpy())


def assign_clusters(args, model_names, clusterings):
    """run_clustering.py:180-272 -- rank r labels shards r::world, one output shard per input shard."""
    from .loader import ShardLoader
    rank, world = _world(args)
    keys = model_key_map(model_names)
    shard_paths = D.expand_shards(args.data.path)[rank::world]
    prefix = '' if args.clustering.cached_epoch is None else 'epoch_{}_'.format(args.clustering.cached_epoch)
    batch_size = max(int(args.data.batch_size), 1024)
    saved_paths = []
    print("extracting clustering for models: {}".format(model_names))
    for shard_path in shard_paths:
        out_path = args.data.output.path / (prefix + shard_path.stem + '.pkl')
        if out_path.is_file():                                                               # :248-250
            continue
        ids = defaultdict(list)
        shards = {name: defaultdict(dict) for name in model_names}
        for batch in ShardLoader([shard_path], batch_size, drop_last=False, workers=0):
            per_model = {name: _extract_batch(batch[keys[name]], clusterings[name]) for name in model_names}
            for j, idx in enumerate(batch['idx']):
                shard_name = batch['shard_name'][j]
                if idx in shards[model_names[0]][shard_name]:                                # dedupe :236
                    continue
                ids[shard_name].append(idx)
                for name in model_names:
                    shards[name][shard_name][idx] = {
                        'assignments': per_model[name][j], 'filename': batch['filename'][j],
                        'shard_name': shard_name, 'shard_size': batch['shard_size'][j], 'idx': idx}
        for shard_name, id_list in ids.items():
            data = [{'model_key': name, 'data': shards[name][shard_name], **MODELS[name]['tag']}
                    for name in model_names]
            saved_paths.append(save_assignments(args, shard_name, id_list, data, prefix=prefix))
    return saved_paths


def run_clustering(args):
    """run_clustering.py:25-29."""
    model_names = filter_models(args)
    with torch.no_grad():
        clusterings = train_clusters(args, model_names)
        return assign_clusters(args, model_names, clusterings)
