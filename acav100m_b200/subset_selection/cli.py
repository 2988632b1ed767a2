"""``python -m acav100m_b200.subset_selection.cli run --shards_path=... --meta_path=... --out_path=...``

Drop-in for ``subset_selection/code/cli.py run`` (reference cli.py:17-104): same flags and defaults
(config.py), same path handling (a bare directory gets ``/output.csv``; metadata defaults to the shard
directory), same append-mode CSV.  ``--measure_name=mem_mi`` selects the CUDA engine; the reference's
default ``batch_mi`` and the chunked / contrastive modes are not built yet and say so.
"""
import copy
import datetime
import sys
import time
from pathlib import Path

from .. import hostio
from .config import defaults
from .run import run_single


def get_args(**kwargs):
    """subset_selection/code/args.py:11-30."""
    args = hostio.update_args(copy.deepcopy(defaults), kwargs)
    hostio.resolve_paths(args, Path('.').resolve())
    args = hostio.objectify(args)
    args.computation.device = 'cuda' if args.computation.use_gpu else 'cpu'
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
    if args.measure_name == 'contrastive' or args.chunk_size is not None:
        raise NotImplementedError("contrastive / chunked selection (reference chunk.py, run_contrastive.py) "
                                  "is outside the CUDA hot path built so far (DESIGN.md scope table)")
    run_single(args)
    print('done. total time elasped: {}'.format(str(datetime.timedelta(seconds=time.time() - start))))


def main(argv=None):
    command, kwargs = hostio.parse_cli(sys.argv[1:] if argv is None else argv)
    if command != 'run':
        raise SystemExit("usage: cli.py run --shards_path=... --meta_path=... --out_path=... [--a.b.c=v ...]")
    run(**kwargs)


if __name__ == '__main__':
    main()
