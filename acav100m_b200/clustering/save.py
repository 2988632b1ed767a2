"""Cluster-shard writer and run log (reference clustering/code/save.py:9-17, 48-74).

Cluster shard ``[epoch_{e}_]shard-NNNNNN.pkl`` = list of rows ``{'filename', 'shard_name', 'shard_size',
'video_assignments': [feat..], 'audio_assignments': [feat..]}`` with ``feat = {'model_key',
'extractor_name', 'dataset', 'array': {'layer_i': np.int64}}`` -- the layout subset_selection's
``format_row`` reads (subset_selection/code/dataloader.py:17-36).
"""
from .. import hostio


def save_assignments(args, shard_name, ids, data, prefix=''):
    """`data`: list of {'model_key', 'name', 'dataset', 'data': {idx: {'assignments': ..., meta..}}}."""
    res = []
    for idx in ids:
        row = {'video_assignments': [], 'audio_assignments': []}
        for model_feat in data:
            point = model_feat['data'][idx]
            array = point['assignments']
            if isinstance(array, (tuple, list)):
                array = {'layer_{}'.format(i): v for i, v in enumerate(array)}
            feature = {'model_key': model_feat['model_key'], 'extractor_name': model_feat['name'],
                       'dataset': model_feat['dataset'], 'array': array}
            for key in ('filename', 'shard_size', 'shard_name'):
                row[key] = point[key]
            if model_feat['model_key'] in args.model_types.audio:
                row['audio_assignments'].append(feature)
            else:
                row['video_assignments'].append(feature)
        res.append(row)
    out_path = args.data.output.path / (prefix + shard_name + '.pkl')
    out_path.parent.mkdir(exist_ok=True, parents=True)
    hostio.dump_pickle(res, out_path)
    return out_path


def store_shards_set(args, saved_paths):
    """``log_{hostname}_{pid}_{timestamp}.json`` listing the shards this run wrote; subset_selection
    groups shards into partitions by these logs (dataloader.py:72-83)."""
    if len(saved_paths) == 0:
        print("All shards already processed")
        return None
    out_path = saved_paths[0].parent / ('log_' + args.run_id + '.json')
    hostio.dump_json({**args.run_info, 'shards': [p.stem for p in saved_paths]}, out_path)
    return out_path
