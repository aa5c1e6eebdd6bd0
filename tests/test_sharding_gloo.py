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


# ----------------------------------------------------------------------------- one long trajectory over several ranks (config 4 sharded)
def _halo_worker(rank, world, port, lens, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo = sum(lens[:rank]); n = lens[rank]
    t = torch.arange(lo, lo + n, dtype=torch.float64)                       # knot k of the whole track has t = k
    xyz = torch.stack([10 * t, 100 * t, 1000 * t], dim=1)
    t_ext, xyz_ext, n_left = sharding.exchange_halo(t, xyz, halo=4)
    # zone record of this rank: mean lon, mean lat, zone, south, valid count (rank 1 of 3 has no valid row)
    zones = {0: [8.4, 49.0, 32.0, 0.0, 100.0], 1: [float("nan"), float("nan"), float("nan"), 0.0, 0.0], 2: [13.1, 49.2, 33.0, 0.0, 300.0]}
    z = torch.tensor(zones.get(rank, zones[0]), dtype=torch.float64)
    zone, south, mine_ok, counts = sharding.global_zone(z)
    assert counts == [int(zones.get(r, zones[0])[4]) for r in range(world)]
    rows = sharding.all_gather_rows(torch.full((3,), float(rank), dtype=torch.float64))
    torch.save({"t": t_ext, "xyz": xyz_ext, "n_left": n_left, "zone": (zone, south, mine_ok), "rows": rows}, os.path.join(out_dir, f"h{rank}.pt"))
    dist.destroy_process_group()


@pytest.mark.parametrize("lens", [[10, 12], [9, 2, 11], [3, 8, 6]])
def test_halo_exchange_and_global_zone(tmp_path, lens):
    """exchange_halo: every rank ends up with its own knots plus up to `halo` knots of each neighbour, in track order, shards
    shorter than the halo included; global_zone: count-weighted means over the ranks, ranks without valid rows ignored."""
    world = len(lens)
    port = _free_port()
    mp.spawn(_halo_worker, args=(world, port, lens, str(tmp_path)), nprocs=world, join=True)
    outs = [torch.load(tmp_path / f"h{r}.pt") for r in range(world)]
    for r, o in enumerate(outs):
        lo = sum(lens[:r]); hi = lo + lens[r]
        left = min(4, lens[r - 1]) if r > 0 else 0
        right = min(4, lens[r + 1]) if r < world - 1 else 0
        want = torch.arange(lo - left, hi + right, dtype=torch.float64)
        assert o["n_left"] == left
        assert torch.equal(o["t"], want)
        assert torch.equal(o["xyz"], torch.stack([10 * want, 100 * want, 1000 * want], dim=1))
        assert torch.equal(o["rows"], torch.arange(world, dtype=torch.float64).reshape(-1, 1).expand(world, 3))
    if world == 3:
        # means: lon (8.4 * 100 + 13.1 * 300) / 400 = 11.925 -> zone 32; rank 2's own block pointed to zone 33
        assert [o["zone"] for o in outs] == [(32, False, True), (32, False, True), (32, False, False)]
    else:
        assert all(o["zone"] == (32, False, True) for o in outs)
