"""N>1 host logic on CPU: world_size-2 (and 3) gloo process groups exercise the trajectory
sharding and the final gather of ATE statistics that bench.py uses on the GPUs with NCCL."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gps_optimize_slam_b200 import sharding


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = sharding.shard_range(total, rank, world)
    # "statistics" of trajectory b = (b, 2b, 3b, rank): rank-order concatenation must restore 0..total-1
    b = torch.arange(lo, hi, dtype=torch.float64)
    local = torch.stack([b, 2 * b, 3 * b, torch.full_like(b, rank)], dim=1)
    full = sharding.gather_stats(local, total_rows=total)
    counts = [sharding.shard_range(total, r, world)[1] - sharding.shard_range(total, r, world)[0] for r in range(world)]
    full2 = sharding.gather_stats(local, total_rows=total, counts=counts)     # one collective, no height exchange
    assert torch.equal(full, full2)
    t = sharding.max_over_ranks(10.0 + rank, "cpu")
    torch.save({"full": full, "t": t, "range": (lo, hi)}, os.path.join(out_dir, f"r{rank}.pt"))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,total", [(2, 4096), (2, 7), (3, 10)])
def test_shard_and_gather(tmp_path, world, total):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, total, str(tmp_path)), nprocs=world, join=True)
    outs = [torch.load(tmp_path / f"r{r}.pt") for r in range(world)]
    covered = []
    for r, o in enumerate(outs):
        lo, hi = o["range"]
        covered += list(range(lo, hi))
        assert o["t"] == 10.0 + world - 1                       # max over ranks
        assert torch.equal(o["full"][:, 0], torch.arange(total, dtype=torch.float64))
        assert torch.equal(o["full"], outs[0]["full"])          # every rank holds the same table
    assert covered == list(range(total))
    sizes = [o["range"][1] - o["range"][0] for o in outs]
    assert max(sizes) - min(sizes) <= 1


def test_shard_range_validation():
    assert sharding.shard_range(1 << 20, 3, 8) == (393216, 524288)
    with pytest.raises(ValueError):
        sharding.shard_range(10, 2, 2)
    t = torch.ones(3, 4)
    assert sharding.gather_stats(t) is t                        # no process group: pass-through
