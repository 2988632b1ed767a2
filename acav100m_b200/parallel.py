"""Host-side multi-GPU protocol of the two operators (one process per GPU, torch.distributed).

Nothing here computes on tensors' values; it only says who owns what and how per-rank results are
combined, so the same code runs under NCCL on GPUs and under gloo in the CPU tests
(tests/test_multirank_cpu.py drives it with the oracle standing in for the CUDA engine).

k-means (reference sgd_clustering.py:94-129, is_distributed branch): replicated centers/counts, each
rank assigns its own slice of the global batch, histograms and deltas are summed over ranks, `count`
advances by the global batch.  The reference's all_gather of the batch (:97) is only used for its
length and is replaced by arithmetic.

greedy MI (new -- the reference's multi-GPU mode is independent chunks, chunk.py:21-53): the
candidate list is cut into `world` contiguous position ranges; every iteration each rank scores its
range and contributes one (key, cell) pair, key = (orderable fp32 score << 32) | (0xFFFFFFFF - global
position); the pair with the largest key -- highest score, earliest position: exactly the reference's
first-arg-max -- is applied by every rank to its replicated table.
"""
import numpy as np


def shard_bounds(n, rank, world):
    """Contiguous range [lo, hi) of a list of n items owned by `rank` (keeps list order global)."""
    return (n * rank) // world, (n * (rank + 1)) // world


def orderable_u32(score):
    """fp32 -> uint32 whose unsigned order equals the float order (device twin: common.cuh orderable)."""
    s = np.float32(score)
    if s == 0:
        s = np.float32(0.0)
    u = int(np.array(s, dtype=np.float32).view(np.uint32))
    return (~u & 0xFFFFFFFF) if (u & 0x80000000) else (u | 0x80000000)


def score_from_orderable(u):
    u = int(u) & 0xFFFFFFFF
    bits = (u & 0x7FFFFFFF) if (u & 0x80000000) else (~u & 0xFFFFFFFF)
    return float(np.array(bits, dtype=np.uint32).view(np.float32))


def pack_key(score, position):
    """0 is reserved for "no candidate left"."""
    return (orderable_u32(score) << 32) | (0xFFFFFFFF - int(position))


def unpack_key(key):
    key = int(key) & 0xFFFFFFFFFFFFFFFF
    return score_from_orderable(key >> 32), 0xFFFFFFFF - (key & 0xFFFFFFFF)


def pack_cell(c1, c2):
    return (int(c1) << 16) | int(c2)


def unpack_cell(cell):
    cell = int(cell)
    return (cell >> 16) & 0xFFFF, cell & 0xFFFF


def record_words(d):
    """uint64 words of a multi-pair winner record (device twin: acav_mi_pairs_record_words): the key, then the
    winner's d cluster ids, 16 bits each, four per word."""
    return 1 + (int(d) + 3) // 4


def pack_record(key, ids):
    """-> list of unsigned ints [key, ids 0..3, ids 4..7, ...] (mip_emit_kernel's layout)."""
    words = [int(key) & 0xFFFFFFFFFFFFFFFF] + [0] * (record_words(len(ids)) - 1)
    for j, c in enumerate(ids):
        words[1 + j // 4] |= (int(c) & 0xFFFF) << (16 * (j % 4))
    return words


def unpack_record(words, d):
    """-> (key, [d ids])."""
    return int(words[0]) & 0xFFFFFFFFFFFFFFFF, [(int(words[1 + j // 4]) >> (16 * (j % 4))) & 0xFFFF for j in range(d)]


def combine_records(records):
    """records: iterable of word lists -> the one with the largest key (None if every key is 0)."""
    best = None
    for rec in records:
        if (int(rec[0]) & 0xFFFFFFFFFFFFFFFF) > (0 if best is None else int(best[0]) & 0xFFFFFFFFFFFFFFFF):
            best = rec
    return best


def combine_pairs(pairs):
    """pairs: iterable of (key, cell) as unsigned ints -> the winning (key, cell) or (0, 0)."""
    best = (0, 0)
    for key, cell in pairs:
        key = int(key) & 0xFFFFFFFFFFFFFFFF
        if key > best[0]:
            best = (key, int(cell))
    return best


def sharded_greedy(engine, dist, world, n_picks, new_pair_buffer, new_gather_buffer, on_pick=None):
    """Greedy loop over sharded candidates.

    engine.local_best(out_pair)         -> fills a 2 x int64 buffer with this rank's (key, cell)
    engine.apply(all_pairs, world, i)   -> every rank applies the winner among the gathered pairs
    The buffers are int64 tensors on the engine's device (bit patterns of uint64); the gather buffer is
    flat [2 * world] = world consecutive (key, cell) pairs (a layout both NCCL and gloo accept).
    The multi-pair engine (P > 1) runs the same loop with records of `record_words(d)` words -- key + the
    winner's d cluster ids -- instead of (key, cell) pairs."""
    mine = new_pair_buffer()
    allp = new_gather_buffer(world)
    for i in range(n_picks):
        engine.local_best(mine)
        dist.all_gather_into_tensor(allp, mine)
        engine.apply(allp, world, i)
        if on_pick is not None:
            on_pick(i, allp)


def kmeans_global_batch(local_b, world):
    """sgd_clustering.py:101,128 -- `count` advances by the gathered batch length."""
    return local_b * world
