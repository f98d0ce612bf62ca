"""CPU tier: the N>1 host logic (contiguous sharding + gather at the mux boundary) with world_size 2 over gloo."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ffmpeg_b200.sharding import shard_range, gather_to_mux


def test_shard_range_partitions():
    for n in (0, 1, 7, 256, 1000, 48960):
        for world in (1, 2, 3, 4, 8):
            got = [shard_range(n, r, world) for r in range(world)]
            assert got[0][0] == 0 and got[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(got, got[1:]))
            sizes = [b - a for a, b in got]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, n_items):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    a, b = shard_range(n_items, rank, world)
    # stand-in for the per-rank conversion: every item is a 6-byte "frame" derived from its global index
    idx = torch.arange(a, b, dtype=torch.int64)
    local = ((idx[:, None] * 7 + torch.arange(6)[None, :]) % 251).to(torch.uint8)
    out = gather_to_mux(local, dst=0)
    if rank == 0:
        idx = torch.arange(0, n_items, dtype=torch.int64)
        exp = ((idx[:, None] * 7 + torch.arange(6)[None, :]) % 251).to(torch.uint8)
        assert torch.equal(out, exp)
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


def test_gather_world2_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    for n_items in (9, 256):
        mp.spawn(_worker, args=(2, port, n_items), nprocs=2, join=True)
        port += 1


def _mux_worker(rank, world, port, items):
    from ffmpeg_b200.sharding import MuxGather
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mux = MuxGather(items, (6,), torch.uint8, "cpu", dst=0)
    for step in range(3):                                     # the object is reused batch after batch
        idx = torch.arange(rank * items, (rank + 1) * items, dtype=torch.int64)
        local = ((idx[:, None] * 7 + torch.arange(6)[None, :] + step) % 251).to(torch.uint8)
        out = mux.start(local)
        mux.wait()
        if rank == 0:
            idx = torch.arange(0, world * items, dtype=torch.int64)
            exp = ((idx[:, None] * 7 + torch.arange(6)[None, :] + step) % 251).to(torch.uint8)
            assert torch.equal(out, exp), step
        else:
            assert out is None
        dist.barrier()
    dist.destroy_process_group()


def test_mux_gather_object_world2_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_mux_worker, args=(2, port, 17), nprocs=2, join=True)


def test_spread_over_numa():
    from ffmpeg_b200.sharding import spread_over_numa
    import pytest
    box = [(i, 0 if i < 4 else 1) for i in range(8)]          # an HGX box: GPUs 0-3 on socket 0, 4-7 on socket 1
    assert spread_over_numa(box, 1) == [0]
    assert spread_over_numa(box, 2) == [0, 4]
    assert spread_over_numa(box, 4) == [0, 4, 1, 5]
    assert spread_over_numa(box, 8) == [0, 4, 1, 5, 2, 6, 3, 7]
    assert spread_over_numa([(0, -1), (1, -1)], 2) == [0, 1]  # no NUMA information: plain order
    assert spread_over_numa([(2, 1), (3, 1), (0, 0)], 3) == [0, 2, 3]
    with pytest.raises(ValueError):
        spread_over_numa(box[:2], 4)
