"""Which clustering columns form contingency tables (reference ``pairing.py:5-41``)."""
from collections import defaultdict
from itertools import combinations, product


def get_cluster_pairing(keys, cluster_pairing):
    kinds = {'diagonal': get_diagonal, 'bipartite': get_bipartite, 'combination': get_combination}
    cluster_pairing = cluster_pairing.lower()
    assert cluster_pairing in kinds, f"invalid cluster pairing type: {cluster_pairing}"
    return kinds[cluster_pairing](keys)


def get_combination(keys):
    return list(combinations(range(len(keys)), 2))


def _group(keys, field):
    groups = defaultdict(list)
    for idx, key in enumerate(keys):
        groups[key[field]].append(idx)
    return list(groups.values())


def get_bipartite(keys):
    """one index group per dataset+model name (key[0]); every cross-group tuple is a pairing"""
    return list(product(*_group(keys, 0)))


def get_diagonal(keys):
    """indices sharing a clustering (layer) name (key[1]) form a pairing"""
    return _group(keys, 1)
