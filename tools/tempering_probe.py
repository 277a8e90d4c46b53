"""Parallel tempering over several GPUs (BASELINE config 4 shape):
    torchrun --nproc-per-node N tools/tempering_probe.py [n] [kind] [steps] [ms] [start]
Every rank owns 4736 replicas on a geometric temperature ladder (per-replica exp(-delta/T) acceptance in the
kernel); after every step the ranks all-gather (cost, temperature), compute the same swap decisions and
only temperatures move; every fourth step the best slab is broadcast over NCCL."""
import lzma
import os
import sys
import time
sys.path.insert(0, '.')
import numpy as np
import torch
import torch.distributed as dist
import megalania_b200 as mg
from megalania_b200.tempering import ReplicaExchange, temperature_ladder
from tools import corpus
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
kind = sys.argv[2] if len(sys.argv) > 2 else "text"
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 12
ms = float(sys.argv[4]) if len(sys.argv) > 4 else 500.0
start = sys.argv[5] if len(sys.argv) > 5 else "greedy"
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
data = corpus.make(kind, n)
ctx = mg.Context(data, device=local)
chains = 4736
an = mg.Annealer(ctx, chains, seed=17 + 1000003 * rank)
init = None
if start == "greedy":
    from oracle import oracle_lib
    init = oracle_lib.Port().greedy_slab(data)
an.set_slab(init, adopt_cost=True)
rex = ReplicaExchange(dist, device=f"cuda:{local}", seed=5)
ladder = temperature_ladder(chains * world, 64.0, 16384.0)     # 1/32 bit .. 8 bits
temps = ladder[rank::world].copy()
buf = torch.empty(n * 8, dtype=torch.uint8, device=f"cuda:{local}")
t0 = time.time()
total = 0
for s in range(steps):
    st = an.run(1_000_000, schedule=mg.SCHEDULE_TEMPERATURE, temperatures=temps, first_eval=mg.CONTINUE_EVALS,
                cycle_budget=int(ms * 1.965e6), suspend=True)
    cur, best = an.costs()
    swapped = rex.exchange(cur.astype(np.int64), temps)
    moved = int((swapped != temps).sum())
    temps = swapped
    if (s + 1) % 4 == 0:
        nz = np.where(best > 0, best, np.iinfo(np.uint64).max)
        rex.broadcast_best(int(nz.min()),
                           lambda b: an.export_slab(int(nz.argmin()), True, b.data_ptr()),
                           lambda b: an.import_slab(int(cur.argmax()), b.data_ptr(), adopt_cost=True), buf)
    ev = torch.tensor([st["evals"], int(np.where(best > 0, best, np.iinfo(np.uint64).max).min())], dtype=torch.int64, device=f"cuda:{local}")
    evs = [torch.empty_like(ev) for _ in range(world)]
    dist.all_gather(evs, ev)
    total += sum(int(e[0]) for e in evs)
    if rank == 0:
        print(s + 1, "t=%.1fs" % (time.time() - t0), "best bytes %.1f" % (min(int(e[1]) for e in evs) / 16384 + 18),
              "evals", total, "temperatures swapped on this rank", moved, flush=True)
cur, best = an.costs()
nz = np.where(best > 0, best, np.iinfo(np.uint64).max)
mine = torch.tensor([int(nz.min())], dtype=torch.int64, device=f"cuda:{local}")
allb = [torch.empty_like(mine) for _ in range(world)]
dist.all_gather(allb, mine)
if rank == int(np.argmin([int(b) for b in allb])):
    stream = ctx.encode_slab(an.get_slab(int(nz.argmin()), best=True))
    assert lzma.decompress(stream, format=lzma.FORMAT_ALONE) == data
    print("final .lzma", len(stream), "bytes from rank", rank, "of", world, "; round-trips", flush=True)
an.close()
dist.barrier()
dist.destroy_process_group()
