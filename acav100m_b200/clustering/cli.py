"""``python -m acav100m_b200.clustering.cli cluster --feature_path=... --out_path=... --meta_path=...``

Drop-in for ``clustering/code/cli.py cluster`` (reference cli.py:8-31, script.py:18-69): same flags
(every dotted ``--a.b.c=v`` override of config.py), same outputs (cluster shards, ``cache_epoch_*``
checkpoints, ``log_*.json``), ends by printing ``done``.  One process per GPU: launched under torchrun
(RANK / WORLD_SIZE in the environment) it joins the NCCL group; with ``--computation.num_gpus=G > 1``
outside torchrun it spawns G ranks itself like the reference's ``torch.multiprocessing.spawn``.
"""
import os
import sys

import torch

from .. import hostio
from .args import cli_aliases, get_args
from .run_clustering import run_clustering
from .save import store_shards_set


def _run_rank(rank, world, kwargs, port):
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", str(port))
    torch.cuda.set_device(rank % max(torch.cuda.device_count(), 1))
    dist.init_process_group("nccl", rank=rank, world_size=world)
    try:
        _cluster(kwargs)
    finally:
        dist.destroy_process_group()


def _cluster(kwargs):
    args = get_args(**kwargs)
    args.data.output.path.mkdir(parents=True, exist_ok=True)
    saved_paths = run_clustering(args)
    store_shards_set(args, saved_paths)
    return saved_paths


def cluster(**kwargs):
    kwargs = cli_aliases(kwargs)
    world_env = int(os.environ.get("WORLD_SIZE", "1"))
    if world_env > 1:                                   # torchrun
        import torch.distributed as dist
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        dist.init_process_group("nccl")
        kwargs.setdefault('computation.num_gpus', world_env)
        try:
            _cluster(kwargs)
        finally:
            dist.destroy_process_group()
    else:
        probe = get_args(**kwargs)
        n_shards = len(hostio.braceexpand(str(probe.data.path)))
        num_gpus = min(probe.computation.num_gpus or 1, max(n_shards, 1))            # script.py:22,37
        if num_gpus > 1:
            kwargs['computation.num_gpus'] = num_gpus
            torch.multiprocessing.spawn(_run_rank, nprocs=num_gpus,
                                        args=(num_gpus, kwargs, probe.computation.master_port))
        else:
            kwargs['computation.num_gpus'] = 1
            _cluster(kwargs)
    print('done')


def main(argv=None):
    command, kwargs = hostio.parse_cli(sys.argv[1:] if argv is None else argv)
    if command not in ('cluster', 'run'):
        raise SystemExit("usage: cli.py cluster --feature_path=... --out_path=... [--a.b.c=v ...]")
    cluster(**kwargs)


if __name__ == '__main__':
    main()
