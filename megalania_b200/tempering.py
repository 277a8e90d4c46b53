"""Multi-GPU host logic: replica exchange (parallel tempering) and best-slab broadcast.

Chains never interact inside an annealing step; between steps the ranks exchange a few bytes:
  * all-gather of every chain's (cost, temperature) -> swap decisions computed identically on
    every rank from a shared counter-based generator -> only TEMPERATURES move, no slab does;
  * all-gather of each rank's best cost -> the arg-min rank broadcasts its best packed slab
    (8*n bytes) -> the other ranks re-seed their worst chain with it.  This mirrors the
    reference restarting every later epoch from `packets_best` (src/main.c:75-77,89-92).
The collectives run through torch.distributed (NCCL over NVLink on GPUs; gloo in the CPU tests).
"""
from __future__ import annotations

import math

import numpy as np

_MASK = (1 << 64) - 1


def _splitmix64(x: int) -> int:
    x = (x + 0x9E3779B97F4A7C15) & _MASK
    z = x
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _MASK
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _MASK
    return z ^ (z >> 31)


def _uniform(seed: int, round_index: int, pair: int) -> float:
    return (_splitmix64(seed ^ (round_index * 0x9E3779B97F4A7C15 + pair * 0xD1342543DE82EF95) & _MASK) >> 11) / float(1 << 53)


def temperature_ladder(replicas: int, t_min: float, t_max: float) -> np.ndarray:
    """Geometric ladder, coldest first (temperatures in 1/2048-bit cost units)."""
    if replicas == 1:
        return np.array([t_min], dtype=np.float32)
    r = (t_max / t_min) ** (1.0 / (replicas - 1))
    return np.array([t_min * r ** i for i in range(replicas)], dtype=np.float32)


def exchange_temperatures(costs: np.ndarray, temps: np.ndarray, round_index: int, seed: int = 0) -> np.ndarray:
    """One round of replica exchange over ALL replicas (global arrays, identical on every rank).

    Replicas adjacent on the temperature ladder (even pairs on even rounds, odd pairs on odd
    rounds) swap temperatures with the Metropolis probability
        min(1, exp((1/T_i - 1/T_j) * (C_i - C_j)))
    Returns the new temperature of every replica.  Pure function of its arguments."""
    costs = np.asarray(costs, dtype=np.float64)
    temps = np.asarray(temps, dtype=np.float64).copy()
    order = np.argsort(temps, kind="stable")  # ladder position -> replica
    start = round_index & 1
    for pair, lo in enumerate(range(start, len(order) - 1, 2)):
        i, j = order[lo], order[lo + 1]
        ti, tj = temps[i], temps[j]
        if ti <= 0 or tj <= 0:
            continue
        delta = (1.0 / ti - 1.0 / tj) * (costs[i] - costs[j])
        if delta >= 0 or _uniform(seed, round_index, pair) < math.exp(max(delta, -700.0)):
            temps[i], temps[j] = tj, ti
    return temps.astype(np.float32)


def arg_best(costs) -> int:
    """Index of the smallest non-zero cost (0 = "no evaluation yet"); ties -> lowest index."""
    best, arg = None, 0
    for i, c in enumerate(costs):
        c = int(c)
        if c != 0 and (best is None or c < best):
            best, arg = c, i
    return arg


class ReplicaExchange:
    """Collectives between annealing steps.  `dist` is torch.distributed (initialised) or None."""

    def __init__(self, dist=None, device=None, seed: int = 0):
        self.dist = dist
        self.device = device
        self.seed = seed
        self.round = 0
        self.rank = dist.get_rank() if dist is not None else 0
        self.world = dist.get_world_size() if dist is not None else 1

    def _all_gather(self, values: np.ndarray, dtype):
        import torch
        if self.dist is None:
            return np.asarray(values)[None, :]
        t = torch.as_tensor(np.asarray(values), dtype=dtype, device=self.device)
        out = [torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(out, t)
        return np.stack([o.cpu().numpy() for o in out])

    def exchange(self, local_costs, local_temps) -> np.ndarray:
        """Replica exchange across all ranks; returns this rank's new temperatures."""
        import torch
        costs = self._all_gather(np.asarray(local_costs, dtype=np.int64), torch.int64)
        temps = self._all_gather(np.asarray(local_temps, dtype=np.float32), torch.float32)
        per_rank = costs.shape[1]
        new = exchange_temperatures(costs.reshape(-1), temps.reshape(-1), self.round, self.seed)
        self.round += 1
        return new.reshape(self.world, per_rank)[self.rank]

    def broadcast_best(self, local_best_cost: int, export_fn, import_fn, buffer) -> tuple[int, int]:
        """All-gather the ranks' best costs; the arg-min rank fills `buffer` (a tensor on
        `device`) through export_fn(buffer) and broadcasts it; every other rank receives it and
        calls import_fn(buffer).  Returns (source rank, global best cost)."""
        import torch
        costs = self._all_gather(np.array([int(local_best_cost)], dtype=np.int64), torch.int64).reshape(-1)
        src = arg_best(costs)
        if self.dist is None:
            return 0, int(costs[0])
        if self.rank == src:
            export_fn(buffer)
        self.dist.broadcast(buffer, src=src)
        # the library works on its own stream: the broadcast must have landed before it reads the buffer
        if getattr(buffer, "is_cuda", False):
            torch.cuda.synchronize(buffer.device)
        if self.rank != src:
            import_fn(buffer)
        return src, int(costs[src])
