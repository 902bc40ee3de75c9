"""Multi-GPU partitioning of the path: one process per GPU, no data-path collective.

Level 1: whole tuner streams -> ranks.  Level 2: one tuner stream on several GPUs -- every rank channelizes the
same input and keeps a contiguous slice of the polyphase bins (and demodulates those).  Level 3: channel-domain
banks -- a contiguous slice of the channel rows per rank.  torch.distributed is only used for the timing barrier
and the max-over-ranks reduction.
"""


def balanced_slice(n_items, rank, world):
    """contiguous [start, stop) of n_items for `rank`; sizes differ by at most one, earlier ranks get the extras"""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError("rank %d outside world of %d" % (rank, world))
    base, extra = divmod(int(n_items), world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def stream_assignment(n_streams, world):
    """level 1: tuner stream indexes per rank"""
    return [list(range(*balanced_slice(n_streams, r, world))) for r in range(world)]


def bin_slice(channel_count, rank, world):
    """level 2: polyphase bins [start, stop) this rank extracts and demodulates"""
    return balanced_slice(channel_count, rank, world)


def channel_slice(n_channels, rank, world):
    """level 3: rows of a channel-domain bank owned by this rank"""
    return balanced_slice(n_channels, rank, world)


def max_over_ranks(values, device=None):
    """element-wise MAX of a list of floats over all ranks (device time of a step = slowest rank)"""
    import torch
    import torch.distributed as dist
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()
