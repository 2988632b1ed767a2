"""CPU restatement of the reference's default CLI measure ``batch_mi`` (stochastic batch greedy).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows ``subset_selection/code/measures/batch.py`` (``EfficientBatchMI`` :10-260) on top of the dense
measure of ``measures/mi.py`` (``init_cache`` :32-39, ``one_hot``/``gather_pairs`` :61-74, ``calc_MI``
:85-91, ``get_last`` :93-98, ``_add_samples`` :127-148).  Every iteration the reference (1) reshuffles ALL
remaining candidates with ``torch.randperm`` on the global CPU generator, (2) scores the first B of them
with the dense fp32 MI of (table + one-hot), (3) keeps the top k, (4) adds them to the table and
(5) re-appends the B-k losers in ascending id order.  Pinned by ``tests/golden/bmi_*.npz``.
"""
import math

import numpy as np
import torch

EPS = np.finfo('float64').eps


def one_hot(x, C):
    """mi.py:68-74."""
    out = torch.zeros((*x.shape, C), dtype=torch.float)
    out.scatter_(-1, x.unsqueeze(-1), torch.ones((*x.shape, 1), dtype=torch.float))
    return out


def sample_tables(assignments, pairs, ids, C):
    """``sample_batch`` batch.py:34-54 / ``_add_samples`` mi.py:127-143 -> dense one-hot tables
    N [m,P,C,C], a [m,P,C] (index c2), b [m,P,C] (index c1), n [m,P]."""
    rows = assignments.index_select(0, ids)                       # [m, D]
    oh = one_hot(rows, C)                                         # [m, D, C]
    pair_ids = torch.as_tensor(pairs, dtype=torch.long)           # [P, 2]
    p1, p2 = oh[:, pair_ids[:, 0]], oh[:, pair_ids[:, 1]]         # [m, P, C]
    N = torch.einsum('wpa,wpb->wpab', p1, p2)
    a, b = N.sum(2), N.sum(3)
    return {'N': N, 'a': a, 'b': b, 'n': b.sum(-1)}


def dense_mi(last):
    """``calc_MI`` mi.py:85-91 -> [m, P]."""
    N = last['N']
    a = last['a'].unsqueeze(2)
    b = last['b'].unsqueeze(3)
    n = last['n'].unsqueeze(-1).unsqueeze(-1)
    return (N / n * (N.log() + n.log() - (a.log() + b.log()))).sum([2, 3])


def dense_emi(last):
    """``EfficientAMI.calc_EMI`` mi.py:217-231 -> [m, P].  (One hypergeometric term per cell, evaluated at the cell's
    own count -- not sklearn's sum over all possible counts; restated as the reference has it.)"""
    N = last['N']
    a = last['a'].unsqueeze(2)
    b = last['b'].unsqueeze(3)
    n = last['n'].unsqueeze(-1).unsqueeze(-1)
    term1 = (N / n * (N.log() + n.log() - (a.log() + b.log())))
    log_term2 = (a + 1).lgamma() + (b + 1).lgamma() + (n - a + 1).lgamma() + (n - b + 1).lgamma() \
        - ((n + 1).lgamma() + (N + 1).lgamma() + (a - N + 1).lgamma() + (b - N + 1).lgamma()
           + (n - a - b + N + 1).lgamma())
    return (term1 * log_term2.exp()).sum([2, 3])


def entropies(last):
    """``calc_entropy`` / ``calc_entropies`` mi.py:233-245 -> (ha [m, P], hb [m, P])."""
    n = last['n'].unsqueeze(-1)
    pa, pb = last['a'] / n, last['b'] / n
    return -(pa * pa.log()).sum(dim=-1), -(pb * pb.log()).sum(dim=-1)


def dense_ami(last, average_method='arithmetic'):
    """``EfficientAMI.calc_AMI`` mi.py:247-262 with ``generalized_mean`` :200-209 and ``ensure_nonzero`` :193-198."""
    mi = dense_mi(last)
    emi = dense_emi(last)
    ha, hb = entropies(last)
    if average_method == 'max':
        normalizer = torch.max(ha, hb)
    elif average_method == 'min':
        normalizer = torch.min(ha, hb)
    else:
        normalizer = (ha + hb) / 2
    denominator = normalizer - emi
    denominator = torch.max(denominator, torch.full(denominator.shape, EPS, dtype=denominator.dtype))
    return (mi - emi) / denominator


def modify_k(k, B, subset_size, dataset_size, keep_unselected):
    """batch.py:173-188."""
    term = B * subset_size / dataset_size
    if k < term and not keep_unselected:
        k = math.ceil(term)
    return k


def greedy_batch_mi(assignments, ncentroids, pairs, candidates, subset_size, start_indices,
                    batch_size=20, selection_size=4, keep_unselected=True):
    """``EfficientBatchMI._run_greedy`` batch.py:202-260 -> (S, GAIN).  The start indices are counted
    into the table but never into S; |S| == subset_size exactly."""
    a = torch.from_numpy(np.asarray(assignments)).to(torch.long)
    C, P = ncentroids, len(pairs)
    N = torch.full((P, C, C), EPS)                                # init_cache mi.py:32-39
    cache = {'N': N, 'a': N.sum(dim=1), 'b': N.sum(dim=2)}
    cache['n'] = cache['a'].sum(dim=-1)
    cand = torch.as_tensor(list(candidates), dtype=torch.long)
    B = batch_size
    k = modify_k(selection_size, B, subset_size, a.shape[0], keep_unselected)
    add = sample_tables(a, pairs, torch.as_tensor(list(start_indices), dtype=torch.long), C)
    for key in cache:                                             # add_samples batch.py:190-193
        cache[key] = cache[key] + add[key].sum(0)
    S, GAIN = [], []
    while len(S) < subset_size:
        cand = cand.index_select(0, torch.randperm(cand.shape[0]))          # shuffle_candidate_ids :29-32
        batch = cand[:B]
        tabs = sample_tables(a, pairs, batch, C)
        last = {key: cache[key].unsqueeze(0) + tabs[key] for key in tabs}   # get_last mi.py:93-98
        scores = dense_mi(last).mean(dim=-1)                                # calc_ids :143-150
        kk = k if scores.shape[0] >= B else math.floor(B / k * scores.shape[0])
        top, ids = scores.topk(k=kk, dim=0)
        win = sample_tables(a, pairs, batch.index_select(0, ids), C)
        for key in cache:                                                   # update_cache :152-154
            cache[key] = cache[key] + win[key].sum(0)
        chosen = batch.index_select(0, ids)
        cand = cand[B:]                                                     # update_candidates :156-165
        if keep_unselected:
            comb = torch.cat((batch, chosen), dim=0)
            uniq, counts = comb.unique(return_counts=True)
            cand = torch.cat((cand, uniq[counts == 1]), dim=0)
        S += chosen.tolist()
        GAIN += top.tolist()
    return S[:subset_size], GAIN


def greedy_dense_mi(assignments, ncentroids, pairs, candidates, subset_size, start_indices, follow=None,
                    measure='mi', average_method='arithmetic', dtype=torch.float32):
    """``EfficientMI.run_greedy`` mi.py:150-192 with the dense ``calc_MI`` (:85-91) -- or, measure='ami', with
    ``EfficientAMI.calc_AMI`` (:247-262) -- -> (S, GAIN, per-iteration score vectors).  The start indices are in S but
    never in the table.  `follow`: optional list of picks to adopt instead of the arg-max (teacher forcing for the GPU
    tests).  dtype=float64 evaluates the same expressions in double precision: the value the reference's fp32 tensors
    approximate (its lgamma differences cancel catastrophically, see tests/test_ami_cpu.py)."""
    a = torch.from_numpy(np.asarray(assignments)).to(torch.long)
    C, P = ncentroids, len(pairs)
    N = torch.full((P, C, C), EPS, dtype=dtype)                   # init_cache mi.py:32-39
    cache = {'N': N, 'a': N.sum(dim=1), 'b': N.sum(dim=2)}
    cache['n'] = cache['a'].sum(dim=-1)
    cand = torch.as_tensor(list(candidates), dtype=torch.long)
    S, GAIN, ALL = list(start_indices), [], []
    for j in range(len(start_indices), subset_size - 1):
        tabs = sample_tables(a, pairs, cand, C)
        last = {key: cache[key].unsqueeze(0) + tabs[key].to(dtype) for key in tabs}   # get_last mi.py:93-98
        per_pair = dense_mi(last) if measure == 'mi' else dense_ami(last, average_method)
        scores = per_pair.mean(dim=-1)                                      # calc_score :76-80
        score, idx = scores.max(dim=0)
        ALL.append((scores.clone(), cand.clone()))
        if follow is not None:
            idx = (cand == follow[len(GAIN)]).nonzero()[0, 0]
        idx = int(idx)
        for key in cache:                                                   # update_cache :100-102
            cache[key] = last[key][idx]
        S.append(int(cand[idx]))
        GAIN.append(float(scores[idx]))
        cand = torch.cat((cand[:idx], cand[idx + 1:]))                      # remove_idx_all :104-125
    return S, GAIN, ALL
