"""Flag handling of the clustering CLI (reference clustering/code/args.py:11-83, cli.py:12-27)."""
import copy
from pathlib import Path

import torch

from .. import hostio
from .config import defaults


def get_args(**kwargs):
    args = hostio.update_args(copy.deepcopy(defaults), kwargs)
    root = Path(args['root']).resolve()
    args['root'] = root
    hostio.resolve_paths(args, root)
    args = hostio.objectify(args)
    if 'computation.num_gpus' not in kwargs and args.computation.num_gpus is None:
        args.computation.num_gpus = torch.cuda.device_count()          # args.py:18-19
    args.run_info = hostio.get_run_info()
    args.run_id = hostio.get_run_id(args.run_info)
    return args


def cli_aliases(kwargs):
    """cli.py:13-20 -- out_path / feature_path / shards_path / meta_path aliases."""
    kwargs = dict(kwargs)
    if 'out_path' in kwargs:
        kwargs['data.output.path'] = kwargs.pop('out_path')
    if 'feature_path' in kwargs:
        kwargs['shards_path'] = kwargs.pop('feature_path')
    if 'shards_path' in kwargs:
        kwargs['data.path'] = kwargs.pop('shards_path')
    if 'meta_path' in kwargs:
        kwargs['data.meta.path'] = kwargs.pop('meta_path')
    return kwargs
