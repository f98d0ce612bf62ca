"""Multi-GPU sharding of a batch: frames / blocks / transforms are independent, so the batch is cut into
contiguous ranges, one per rank (one process per GPU), with no collective on the data path.  The only exchange is the
optional gather of the finished outputs to the rank that writes the stream (the "mux boundary"): torch.distributed
gather over NCCL/NVLink on GPUs, gloo in the CPU tests."""
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world):
    """Contiguous partition: rank r gets [start, stop) with sizes differing by at most one, in rank order."""
    if world <= 0 or not (0 <= rank < world) or n_items < 0:
        raise ValueError("bad shard request")
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def gather_to_mux(local, dst=0, group=None):
    """Gather per-rank output tensors (same dtype, first dimension = items of that rank, possibly ragged) on rank `dst`
    in rank order.  Returns the concatenated tensor on `dst`, None elsewhere."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c.item()) for c in counts]
    mx = max(counts)
    pad = local
    if local.shape[0] < mx:
        pad = torch.cat([local, local.new_zeros((mx - local.shape[0],) + tuple(local.shape[1:]))])
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad.contiguous(), bufs, dst=dst, group=group)
    if rank != dst:
        return None
    return torch.cat([b[:c] for b, c in zip(bufs, counts)])
