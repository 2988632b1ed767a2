"""``python -m acav100m_b200.subset_selection.cli run --shards_path=... --meta_path=... --out_path=...``

Drop-in for ``subset_selection/code/cli.py run`` (reference cli.py:17-104): same flags and defaults
(config.py), same path handling (a bare directory gets ``/output.csv``; metadata defaults to the shard
directory), same append-mode CSV.  ``--measure_name=mem_mi`` selects the exact greedy CUDA engine, the
default ``batch_mi`` the reference's batched variant (scoring on the device); ``--chunk_size=n`` runs the
reference's chunked multi-GPU mode (chunk.py) and ``reduce`` / ``reduce_csvs`` / ``reduce_pkls`` merge its
caches, ``compare_measures`` the reference's measure comparison (tests.py).  The contrastive baseline
(run_contrastive.py) is out of scope and says so.
"""
import copy
import datetime
import sys
import time
from pathlib import Path

from .. import hostio
from .chunk import reduce_all_pkls, run_chunks
from .compare import compare_measures as _compare_measures
from .config import defaults
from .run import run_single
from .save import merge_all_csvs


def get_args(**kwargs):
    """subset_selection/code/args.py:11-30."""
    args = hostio.update_args(copy.deepcopy(defaults), kwargs)
    hostio.resolve_paths(args, Path('.').resolve())
    args = hostio.objectify(args)
    args.computation.device = 'cuda' if args.computation.use_gpu else 'cpu'
    # args.py:25-30 -- num_gpus defaults to "all", clamped to the devices present (at least one process)
    import torch
    want = args.computation.num_gpus
    have = torch.cuda.device_count() if torch.cuda.is_available() else 0
    args.computation.num_gpus = max(1, min(sys.maxsize if want is None else int(want), have))
    return args


def prepare(**kwargs):
    """cli.py:18-43."""
    args = get_args(**{k: v for k, v in kwargs.items() if k not in ('out_path', 'shards_path', 'meta_path')})
    if 'out_path' in kwargs:
        args.data.output.path = Path(kwargs['out_path'])
    opath = Path(args.data.output.path)
    if opath.stem == opath.name:                    # potential dir
        opath = opath / 'output.csv'
    opath.parent.mkdir(parents=True, exist_ok=True)
    args.data.output.path = opath
    if 'shards_path' in kwargs:
        args.data.path = Path(kwargs['shards_path'])
    if 'meta_path' in kwargs:
        args.data.meta.path = Path(kwargs['meta_path'])
    mpath = args.data.meta.path
    if mpath is None:
        mpath = Path(args.data.path).parent
    mpath = Path(mpath)
    if not mpath.is_dir() and mpath.parent.is_dir():
        mpath = mpath.parent
    args.data.meta.path = mpath
    return args


def run(**kwargs):
    start = time.time()
    args = prepare(**kwargs)
    if args.measure_name == 'contrastive':
        raise NotImplementedError("the contrastive baseline (reference run_contrastive.py) is outside the scope "
                                  "of this implementation (DESIGN.md scope table)")
    if args.chunk_size is None:
        run_single(args)                                   # cli.py:94-104
    else:
        run_chunks(args)
    print('done. total time elasped: {}'.format(str(datetime.timedelta(seconds=time.time() - start))))


def _timed(fn, **kwargs):
    start = time.time()
    args = prepare(**kwargs)
    fn(args)
    print('done. total time elasped: {}'.format(str(datetime.timedelta(seconds=time.time() - start))))


def reduce_csvs(**kwargs):
    """cli.py:53-59."""
    _timed(merge_all_csvs, **kwargs)


def reduce_pkls(**kwargs):
    """cli.py:61-67."""
    _timed(reduce_all_pkls, **kwargs)


def reduce(**kwargs):
    """cli.py:69-78."""
    _timed(lambda args: merge_all_csvs(args) if args.save_cache_as_csvs else reduce_all_pkls(args), **kwargs)


def compare_measures(**kwargs):
    """cli.py:80-83 (``--measure_names="['mem_mi','mi']"`` by default)."""
    report = _compare_measures(prepare(**kwargs))
    print('done')
    return report


def main(argv=None):
    command, kwargs = hostio.parse_cli(sys.argv[1:] if argv is None else argv)
    commands = {'run': run, 'reduce': reduce, 'reduce_csvs': reduce_csvs, 'reduce_pkls': reduce_pkls,
                'compare_measures': compare_measures}
    if command not in commands:
        raise SystemExit("usage: cli.py run|reduce|reduce_csvs|reduce_pkls|compare_measures --shards_path=... "
                         "--meta_path=... --out_path=... [--a.b.c=v ...]")
    commands[command](**kwargs)


if __name__ == '__main__':
    main()
