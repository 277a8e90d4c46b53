"""Cooperative annealing over byte regions: how thousands of chains work on ONE slab.

The reference anneals a single slab sequentially (src/main.c:64-105).  Thousands of independent
chains explore thousands of trajectories, but only the best of them survives a restart, so the
population advances at the pace of one chain.  Edits far apart in the file barely interact (they
meet only through the adaptive model), which makes the search separable by position:

  round:  every chain starts from the same slab and only mutates packets that start inside its own
          byte region (mg_anneal_run_params.regions); `group` chains share a region;
  merge:  for every region the chain of its group with the lowest total cost is taken, the regions
          are stitched into one slab, the seams are repaired with the reference's own repair rule
          and the result is priced exactly (mg_anneal_merge_regions);
  keep:   the merged slab if it is the cheapest slab seen, else the best single chain;
  shift:  region boundaries move every round, so no position stays a seam.

Every cost is exact (the merged slab is re-priced, never estimated) and every slab handed back
decodes to the input.  Host logic only: the hot path is the same annealing kernel.
"""
from __future__ import annotations

import time

import numpy as np

from . import api


def region_plan(n: int, chains: int, group: int, shift: int, min_region: int = 64):
    """bounds [R+1] and the region index of every chain.  R = chains // group regions of equal size
    (at least min_region bytes), the boundaries rotated by `shift` bytes."""
    group = max(1, min(group, chains))
    nreg = max(1, min(chains // group, n // max(1, min_region)))
    size = n / nreg
    off = shift % max(1, int(size))
    cuts = [0] + [min(n - 1, max(1, int(round(off + r * size)))) for r in range(1, nreg)] + [n]
    bounds = np.array(sorted(set(cuts)), dtype=np.uint32)
    nreg = bounds.size - 1
    region_of_chain = np.arange(chains, dtype=np.int64) % nreg
    return bounds, region_of_chain


class CooperativeAnnealer:
    def __init__(self, annealer: api.Annealer, *, group: int = 8, min_region: int = 64, seed: int = 0):
        self.an = annealer
        self.n = annealer.ctx.n
        self.group = group
        self.min_region = min_region
        self.round_no = 0
        self.rng = np.random.default_rng(seed)
        self.best_cost = None
        self.history = []

    def start(self, slab=None) -> int:
        """All chains from `slab` (None = all literals, like packet_slab_new)."""
        self.an.set_slab(slab, adopt_cost=True, reset_best=True)
        cur, _ = self.an.costs()
        self.best_cost = int(cur[0])
        return self.best_cost

    def round(self, *, cycle_budget: int = 0, packet_budget: int = 0, evals: int = 1_000_000, temperature: float = 0.0,
              ladder: bool = False) -> dict:
        """One round: confined annealing, merge, keep the better of {merged, best chain, previous}.
        temperature (1/2048 bit units) > 0 lets chains climb: exp(-delta/T) acceptance.  With `ladder` the
        chains of a region's group run at 0, T/(g-1), 2T/(g-1), ... T: cold ones exploit, hot ones explore."""
        an, n = self.an, self.n
        shift = int(self.rng.integers(0, n))
        bounds, region_of_chain = region_plan(n, an.chains, self.group, shift, self.min_region)
        regions = np.stack([bounds[region_of_chain], bounds[region_of_chain + 1]], axis=1).astype(np.uint32)
        temps = np.full(an.chains, temperature, dtype=np.float32)
        if ladder:
            nreg_ = bounds.size - 1
            member = np.arange(an.chains) // nreg_
            temps = (temperature * member / max(1, member.max())).astype(np.float32)
        t0 = time.perf_counter()
        st = an.run(evals, schedule=api.SCHEDULE_TEMPERATURE, temperatures=temps, first_eval=api.CONTINUE_EVALS,
                    cycle_budget=cycle_budget, packet_budget=packet_budget, suspend=False, regions=regions)
        t1 = time.perf_counter()
        cur, _ = an.costs()
        cur = cur.astype(np.int64)
        nreg = bounds.size - 1
        # per region the cheapest chain of its group (chains c, c + nreg, c + 2 nreg, ... share region c % nreg)
        order = np.lexsort((cur, region_of_chain))
        first = np.searchsorted(region_of_chain[order], np.arange(nreg))
        owners = order[first].astype(np.uint32)
        best_chain = int(np.argmin(cur))
        best_single = int(cur[best_chain])
        # the merged slab goes to the most expensive chain's slot, so the best single chain survives the merge
        dst = int(np.argmax(cur))
        if dst == best_chain:
            dst = (best_chain + 1) % an.chains
        merged = an.merge_regions(bounds, owners, dst_chain=dst)
        t2 = time.perf_counter()
        if merged <= best_single:
            winner, cost, kind = dst, merged, "merged"
        else:
            winner, cost, kind = best_chain, best_single, "single"
        if self.best_cost is not None and cost > self.best_cost and temperature == 0.0:
            kind += "(no gain)"
        an.broadcast_chain(winner)
        t3 = time.perf_counter()
        self.best_cost = cost if self.best_cost is None else min(self.best_cost, cost)
        self.round_no += 1
        rec = {"round": self.round_no, "cost": cost, "merged": merged, "best_single": best_single, "kept": kind,
               "regions": int(nreg), "evals": st["evals"], "kernel_ms": st["kernel_ms"],
               "run_s": t1 - t0, "merge_s": t2 - t1, "broadcast_s": t3 - t2}
        self.history.append(rec)
        return rec

    def slab(self) -> np.ndarray:
        """The current common slab (all chains hold it after a round)."""
        return self.an.get_slab(0)


def distributed_plan(n: int, chains_per_rank: int, world: int, group: int, shift: int, min_region: int = 64):
    """Regions for `world` processes: region r belongs to rank r % world (every rank's regions are spread over
    the whole file), and the chains of a rank are dealt round-robin over that rank's regions.
    Returns (bounds, owner_rank [R], local_region_of_chain(rank) -> global region index per chain)."""
    group = max(1, min(group, chains_per_rank))
    per_rank = max(1, min(chains_per_rank // group, n // max(1, min_region) // max(1, world)))
    bounds, _ = region_plan(n, per_rank * world, 1, shift, min_region)
    nreg = bounds.size - 1
    owner_rank = np.arange(nreg, dtype=np.int64) % world

    def regions_of_chains(rank: int) -> np.ndarray:
        mine = np.nonzero(owner_rank == rank)[0]
        if mine.size == 0:  # more ranks than regions: share region 0's neighbourhood
            mine = np.array([rank % nreg])
        return mine[np.arange(chains_per_rank) % mine.size]

    return bounds, owner_rank, regions_of_chains


class DistributedCooperativeAnnealer(CooperativeAnnealer):
    """The cooperative search over several GPUs (one process per GPU, torch.distributed).  Every rank anneals
    its own regions; the parts are summed over the ranks (all-reduce: disjoint slots), every rank repairs and
    prices the whole - deterministically, so all ranks hold the same slab and cost without a broadcast.  Only
    when a single chain beats the merged slab does a slab travel (from the rank that owns that chain)."""

    def __init__(self, annealer: api.Annealer, dist, device, *, group: int = 8, min_region: int = 64, seed: int = 0):
        super().__init__(annealer, group=group, min_region=min_region, seed=seed)
        import torch
        self.dist = dist
        self.torch = torch
        self.rank = dist.get_rank()
        self.world = dist.get_world_size()
        self.slab_buf = torch.zeros(self.n, dtype=torch.int64, device=device)   # packed slots, summed as integers
        self.abs_buf = torch.zeros(self.n, dtype=torch.int32, device=device)

    def round(self, *, cycle_budget: int = 0, packet_budget: int = 0, evals: int = 1_000_000, temperature: float = 0.0,
              ladder: bool = False) -> dict:
        an, n, torch, dist = self.an, self.n, self.torch, self.dist
        shift = int(self.rng.integers(0, n))  # same generator state on every rank
        bounds, owner_rank, regions_of_chains = distributed_plan(n, an.chains, self.world, self.group, shift, self.min_region)
        mine = regions_of_chains(self.rank)
        regions = np.stack([bounds[mine], bounds[mine + 1]], axis=1).astype(np.uint32)
        temps = np.full(an.chains, temperature, dtype=np.float32)
        t0 = time.perf_counter()
        st = an.run(evals, schedule=api.SCHEDULE_TEMPERATURE, temperatures=temps, first_eval=api.CONTINUE_EVALS,
                    cycle_budget=cycle_budget, packet_budget=packet_budget, suspend=False, regions=regions)
        t1 = time.perf_counter()
        cur, _ = an.costs()
        cur = cur.astype(np.int64)
        nreg = bounds.size - 1
        owners = np.full(nreg, api.NO_OWNER, dtype=np.uint32)
        order = np.lexsort((cur, mine))
        first = np.searchsorted(mine[order], np.unique(mine))
        owners[np.unique(mine)] = order[first].astype(np.uint32)
        owners[owner_rank != self.rank] = api.NO_OWNER
        an.merge_export(bounds, owners, self.slab_buf.data_ptr(), self.abs_buf.data_ptr())
        torch.cuda.synchronize(self.slab_buf.device)
        dist.all_reduce(self.slab_buf, op=dist.ReduceOp.SUM)
        dist.all_reduce(self.abs_buf, op=dist.ReduceOp.SUM)
        best_chain = int(np.argmin(cur))
        best_single = torch.tensor([int(cur[best_chain])], dtype=torch.int64, device=self.slab_buf.device)
        all_best = [torch.empty_like(best_single) for _ in range(self.world)]
        dist.all_gather(all_best, best_single)
        torch.cuda.synchronize(self.slab_buf.device)
        all_best = [int(t.item()) for t in all_best]
        dst = int(np.argmax(cur))
        if dst == best_chain:
            dst = (best_chain + 1) % an.chains
        merged = an.merge_import(self.slab_buf.data_ptr(), self.abs_buf.data_ptr(), dst_chain=dst)
        t2 = time.perf_counter()
        global_best = min(all_best)
        if merged <= global_best:
            an.broadcast_chain(dst)
            cost, kind = merged, "merged"
        else:
            # a single chain somewhere beats the merge: its slab travels from the rank that owns it
            src_rank = int(np.argmin(all_best))
            if self.rank == src_rank:
                an.export_slab(best_chain, False, self.slab_buf.data_ptr())
            dist.broadcast(self.slab_buf, src=src_rank)
            torch.cuda.synchronize(self.slab_buf.device)
            target = best_chain if self.rank == src_rank else dst
            if self.rank != src_rank:
                an.import_slab(target, self.slab_buf.data_ptr(), adopt_cost=True)
            an.broadcast_chain(target)
            cost, kind = global_best, f"single(rank {src_rank})"
        t3 = time.perf_counter()
        self.best_cost = cost if self.best_cost is None else min(self.best_cost, cost)
        self.round_no += 1
        total = torch.tensor([st["evals"]], dtype=torch.int64, device=self.slab_buf.device)
        dist.all_reduce(total, op=dist.ReduceOp.SUM)
        rec = {"round": self.round_no, "cost": cost, "merged": merged, "best_single": global_best, "kept": kind,
               "regions": int(nreg), "evals": int(total.item()), "kernel_ms": st["kernel_ms"],
               "run_s": t1 - t0, "merge_s": t2 - t1, "broadcast_s": t3 - t2}
        self.history.append(rec)
        return rec
