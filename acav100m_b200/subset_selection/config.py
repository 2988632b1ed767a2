"""Default flag tree of ``subset_selection/code/cli.py run`` (reference subset_selection/code/config.py:1-53).
``clustering.columns`` is the only addition: an optional list of two ``(model_key, layer)`` tuples that
restricts the selection to one audio-visual pair when the cluster shards hold more clusterings (the
CUDA engine scores one contingency table; with all ten layer clusterings the reference forms 45)."""

defaults = {
    'data': {
        'path': 'data',
        'output': {'path': 'output.csv'},
        'meta': {'path': None},
    },
    'computation': {
        'random_seed': 0,
        'num_workers': 40,
        'use_gpu': True,
        'master_port': 6105,
        'dist_backend': 'nccl',
        'dist_init_method': 'tcp://localhost:9967',
        'shard_id': 0,
        'num_shards': 1,
        'use_distributed': True,
        'load_async': False,
        'multiprocess_meta_loading': True,
    },
    'subset': {'ratio': 0.2, 'size': None},
    'clustering': {'pairing': 'combination', 'columns': None},
    'batch': {'batch_size': 20, 'selection_size': 4, 'keep_unselected': True},
    'contrastive': {
        'num_epochs': 3, 'num_warmup_steps': 1, 'base_lr': 2e-4, 'train_batch_size': 128,
        'test_batch_size': 128, 'cached_epoch': None, 'train_from_cached': False,
    },
    'measure_name': 'batch_mi',
    'shuffle_candidates': True,
    'chunk_size': None,
    'save_cache_as_csvs': True,
    'log_every': 1000,
    'log_times': 10,
    'verbose': True,
    'debug': False,
}
