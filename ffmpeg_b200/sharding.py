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


class MuxGather:
    """The mux-boundary collect as a reusable object: rank `dst` owns one receive buffer laid out as the final stream
    ([world * items, ...] in rank order), every other rank sends its finished batch straight into its slot (NCCL send / recv over
    NVLink, grouped into one launch), on a side stream so that the next batch's conversion overlaps the transfer.  `dst`'s own
    batch is copied device-to-device into its slot.  Returns nothing on the other ranks."""

    def __init__(self, items, item_shape, dtype, device, dst=0, group=None):
        self.group, self.dst = group, dst
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.items = items
        self.out = None
        if self.rank == dst:
            self.out = torch.empty((self.world * items,) + tuple(item_shape), dtype=dtype, device=device)
        self.stream = torch.cuda.Stream(device=device) if torch.device(device).type == "cuda" else None
        self.done = None

    def start(self, local, after=None):
        """Begin collecting `local` ([items, ...]); `after`: a CUDA event the side stream waits on (the conversion that fills `local`)."""
        assert local.shape[0] == self.items
        ctx = torch.cuda.stream(self.stream) if self.stream is not None else _Null()
        with ctx:
            if self.stream is not None and after is not None:
                self.stream.wait_event(after)
            if self.rank == self.dst:
                ops = []
                for r in range(self.world):
                    slot = self.out[r * self.items:(r + 1) * self.items]
                    if r == self.rank:
                        slot.copy_(local, non_blocking=True)
                    else:
                        ops.append(dist.P2POp(dist.irecv, slot, r, self.group))
            else:
                ops = [dist.P2POp(dist.isend, local, self.dst, self.group)]
            works = dist.batch_isend_irecv(ops) if ops else []
            for w in works:
                w.wait()                         # NCCL: stream-ordered on the side stream, does not block the host
            if self.stream is not None:
                self.done = torch.cuda.Event()
                self.done.record(self.stream)
        return self.out

    def wait(self, stream=None):
        """Make `stream` (default: the current one) wait for the collect started last."""
        if self.done is not None:
            (stream or torch.cuda.current_stream()).wait_event(self.done)


class _Null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def spread_over_numa(gpus, world):
    """Pick `world` GPUs so that ranks alternate between the host's NUMA nodes (each socket feeds its own PCIe root: end-to-end
    runs with host buffers then use both sockets' memory and PCIe bandwidth even at N = 2 or 4).  `gpus`: list of (index, numa_node)
    of the usable GPUs.  Returns the device index for each rank; GPUs of one node keep their order.  With fewer usable GPUs than
    ranks: ValueError."""
    if world > len(gpus):
        raise ValueError("not enough GPUs")
    by_node = {}
    for idx, node in sorted(gpus):
        by_node.setdefault(node, []).append(idx)
    nodes = sorted(by_node)
    order, k = [], 0
    while len(order) < len(gpus):
        for n in nodes:
            if k < len(by_node[n]):
                order.append(by_node[n][k])
        k += 1
    return order[:world]
