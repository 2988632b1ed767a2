"""Single-process driver of the selection stage (reference subset_selection/code/run.py:6-48)."""
from .dataloader import load_data, preprocess
from .run_greedy import run_greedy
from .save import save_output


def run_partition(args, data, metas):
    assignments, shard_names, filenames, clustering_types = preprocess(data, args.clustering.columns)
    return run_greedy(args, assignments, shard_names, filenames, clustering_types,
                      args.subset.size, args.subset.ratio, measure_name=args.measure_name,
                      cluster_pairing=args.clustering.pairing, shuffle_candidates=args.shuffle_candidates,
                      verbose=args.verbose)


def run_single(args):
    partitions, metas = load_data(args.data.path, args.data.meta.path, args.verbose)
    counts, out_path = 0, None
    for k in sorted(partitions.keys()):
        print('running partition {}/{}'.format(k, len(partitions)))
        samples = run_partition(args, partitions[k], metas)
        out_path, count = save_output(samples, metas, args.data.output.path)
        counts += count
    if out_path is None:
        print("No files saved")
    elif args.verbose:
        print("Saved Results: added {} lines to {}".format(counts, out_path))
    return out_path, counts
