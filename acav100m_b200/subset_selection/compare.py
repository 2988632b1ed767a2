"""``cli.py compare_measures`` (reference subset_selection/code/tests.py:10-51): run several measures on every partition
and print how far their selections and score sequences agree.  The reference's version stops in a debugger after
each comparison and calls ``_run_greedy`` without its first argument; this one runs."""
from itertools import combinations

import numpy as np

from .dataloader import load_data, preprocess
from .run_greedy import _run_greedy


def compare_partition(args, data, measure_names):
    """tests.py:23-51 -> ({measure: S}, {measure: GAIN}, [(k1, k2, S equivalence, mean |GAIN diff|)])."""
    assignments, shard_names, filenames, clustering_types = preprocess(data, args.clustering.columns)
    Ss, GAINs = {}, {}
    for name in measure_names:
        S, GAIN, _ = _run_greedy(args, assignments, clustering_types, args.subset.size, args.subset.ratio,
                                 measure_name=name, cluster_pairing=args.clustering.pairing,
                                 shuffle_candidates=args.shuffle_candidates, verbose=args.verbose)
        Ss[name], GAINs[name] = S, GAIN
    report = []
    for k1, k2 in combinations(list(Ss.keys()), 2):
        sames = np.array([int(v1 == v2) for v1, v2 in zip(Ss[k1], Ss[k2])])
        gain_diffs = np.array([abs(v1 - v2) for v1, v2 in zip(GAINs[k1], GAINs[k2])])
        print(k1, 'vs.', k2)
        print('S equivalence: ', sames.mean())
        print('GAIN diff mean: ', gain_diffs.mean())
        report.append((k1, k2, float(sames.mean()), float(gain_diffs.mean())))
    return Ss, GAINs, report


def compare_measures(args):
    """tests.py:10-20; ``args.measure_names`` defaults to ['mem_mi', 'mi'] as in the reference."""
    names = getattr(args, 'measure_names', None) or ['mem_mi', 'mi']
    partitions, metas = load_data(args.data.path, args.data.meta.path, args.verbose)
    return [compare_partition(args, partitions[k], list(names))[2] for k in sorted(partitions.keys())]
