"""Multi-process host logic (world_size 2, gloo, CPU): replica exchange decisions agree on every
rank, temperatures are conserved, the best slab travels from the arg-min rank to the others."""
import os
import socket

import numpy as np
import pytest

from megalania_b200 import tempering


def test_ladder_and_exchange_are_pure():
    lad = tempering.temperature_ladder(8, 100.0, 10000.0)
    assert len(lad) == 8 and abs(lad[0] - 100) < 1e-3 and abs(lad[-1] - 10000) < 1
    assert (np.diff(lad) > 0).all()
    costs = np.array([500, 400, 300, 200, 900, 100, 50, 10])
    a = tempering.exchange_temperatures(costs, lad, 0, seed=3)
    b = tempering.exchange_temperatures(costs, lad, 0, seed=3)
    assert (a == b).all()
    assert sorted(a.tolist()) == sorted(lad.tolist())
    # a hotter replica holding a better (lower) cost always hands its slab the colder temperature
    a = tempering.exchange_temperatures(np.array([1000, 10]), np.array([1.0, 2.0], dtype=np.float32), 0)
    assert a.tolist() == [2.0, 1.0]
    assert tempering.arg_best([0, 7, 3, 0, 3]) == 2
    assert tempering.arg_best([0, 0]) == 0


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ex = tempering.ReplicaExchange(dist, device="cpu", seed=11)
        per_rank = 4
        lad = tempering.temperature_ladder(world * per_rank, 50.0, 5000.0)
        temps = lad[rank * per_rank:(rank + 1) * per_rank]
        rng = np.random.default_rng(100 + rank)
        history = []
        for _ in range(6):
            costs = rng.integers(1000, 2000, per_rank)
            temps = ex.exchange(costs, temps)
            history.append(temps.copy())
        # best slab: rank 1 holds the best cost
        n = 64
        mine = np.full(n * 8, rank + 1, dtype=np.uint8)
        received = {}
        buf = torch.zeros(n * 8, dtype=torch.uint8)
        src, best = ex.broadcast_best(900 - 100 * rank, lambda b: b.copy_(torch.from_numpy(mine)),
                                      lambda b: received.setdefault("slab", b.numpy().copy()), buf)
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), history=np.stack(history), src=src, best=best,
                 got=received.get("slab", np.zeros(0, dtype=np.uint8)))
    finally:
        dist.destroy_process_group()


def test_world_size_two_gloo(tmp_path):
    import torch.multiprocessing as mp
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r0 = np.load(tmp_path / "rank0.npz")
    r1 = np.load(tmp_path / "rank1.npz")
    lad = tempering.temperature_ladder(8, 50.0, 5000.0)
    for step in range(6):
        both = np.concatenate([r0["history"][step], r1["history"][step]])
        assert np.allclose(np.sort(both), np.sort(lad)), "temperatures must be conserved across ranks"
    assert int(r0["src"]) == int(r1["src"]) == 1
    assert int(r0["best"]) == int(r1["best"]) == 800
    assert (r0["got"] == 2).all() and r0["got"].size == 512  # rank 0 received rank 1's slab
    assert r1["got"].size == 0                               # the source keeps its own


def _merge_worker(rank, world, port, out_dir):
    """The exchange step of the multi-GPU cooperative merge, on CPU tensors: every rank fills the slots of
    the regions it owns, the parts are summed (gloo all-reduce) and must cover every slot exactly once."""
    import torch
    import torch.distributed as dist
    from megalania_b200.cooperative import distributed_plan
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n, chains = 5000, 24
        rng = np.random.default_rng(1)  # the same generator state on every rank, as in the driver
        parts = []
        for _ in range(3):
            shift = int(rng.integers(0, n))
            bounds, owner_rank, regions_of_chains = distributed_plan(n, chains, world, 4, shift)
            slots = torch.zeros(n, dtype=torch.int64)
            count = torch.zeros(n, dtype=torch.int64)
            for r in np.unique(regions_of_chains(rank)):
                assert owner_rank[r] == rank
                slots[bounds[r]:bounds[r + 1]] = (rank + 1) * 1000 + int(r)   # stands for the winner's packed slots
                count[bounds[r]:bounds[r + 1]] = 1
            dist.all_reduce(slots, op=dist.ReduceOp.SUM)
            dist.all_reduce(count, op=dist.ReduceOp.SUM)
            expect = np.zeros(n, dtype=np.int64)
            for r in range(bounds.size - 1):
                expect[bounds[r]:bounds[r + 1]] = (int(owner_rank[r]) + 1) * 1000 + r
            parts.append((slots.numpy().copy(), count.numpy().copy(), expect))
        np.savez(os.path.join(out_dir, f"merge{rank}.npz"), slots=np.stack([p[0] for p in parts]),
                 count=np.stack([p[1] for p in parts]), expect=np.stack([p[2] for p in parts]))
    finally:
        dist.destroy_process_group()


def test_cooperative_merge_exchange_gloo(tmp_path):
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_merge_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r0 = np.load(tmp_path / "merge0.npz")
    r1 = np.load(tmp_path / "merge1.npz")
    assert (r0["count"] == 1).all() and (r1["count"] == 1).all()      # disjoint parts, nothing left out
    assert (r0["slots"] == r0["expect"]).all()                          # the sum IS the merged whole
    assert (r0["slots"] == r1["slots"]).all()                           # and every rank holds the same
