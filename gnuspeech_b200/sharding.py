"""Utterance sharding across the GPUs of one box.  Utterances are independent, so a rank only needs to know
which slice of the batch is its own: there is no data-path collective."""


def shard_bounds(n_utterances, world_size):
    """Contiguous, balanced [start, end) per rank; sizes differ by at most one."""
    base, extra = divmod(int(n_utterances), int(world_size))
    bounds, at = [], 0
    for r in range(world_size):
        size = base + (1 if r < extra else 0)
        bounds.append((at, at + size))
        at += size
    return bounds


def shard_range(n_utterances, rank, world_size):
    return shard_bounds(n_utterances, world_size)[rank]


def weak_scaling_first_index(rank, utterances_per_gpu):
    """Weak scaling (bench.py): rank r synthesizes utterances [r*U, (r+1)*U) of the global stream."""
    return int(rank) * int(utterances_per_gpu)
