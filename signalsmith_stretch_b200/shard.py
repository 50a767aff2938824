"""Multi-GPU sharding of a batch of streams.

Streams are independent (all state is per object in the reference), so a batch shards
contiguously over ranks with no data-path collective.  The only exchange is one all-reduce of the
processed-sample counter (and a max of the per-rank device time) for the throughput report.
"""


def shard_range(batch, rank, world):
    """Contiguous range [lo, hi) of stream indices owned by `rank`; sizes differ by at most 1."""
    base, rem = divmod(batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def reduce_throughput(samples_done, seconds, dist=None, device=None):
    """(total samples over all ranks, max seconds over ranks).  `dist` = torch.distributed or None."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return int(samples_done), float(seconds)
    import torch

    dev = device if device is not None else "cpu"
    cnt = torch.tensor([int(samples_done)], dtype=torch.int64, device=dev)
    tmax = torch.tensor([float(seconds)], dtype=torch.float64, device=dev)
    dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    return int(cnt.item()), float(tmax.item())
