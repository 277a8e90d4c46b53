"""Cooperative search over several GPUs:  torchrun --nproc-per-node N tools/coop_dist_probe.py [n] [kind] [rounds] [ms]
Checks on every rank that all ranks hold the same slab and cost after every round."""
import hashlib
import lzma
import os
import sys
import time
sys.path.insert(0, '.')
import numpy as np
import torch
import torch.distributed as dist
import megalania_b200 as mg
from megalania_b200.cooperative import DistributedCooperativeAnnealer
from tools import corpus
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
kind = sys.argv[2] if len(sys.argv) > 2 else "text"
rounds = int(sys.argv[3]) if len(sys.argv) > 3 else 10
ms = float(sys.argv[4]) if len(sys.argv) > 4 else 250.0
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
data = corpus.make(kind, n)
ctx = mg.Context(data, device=local)
an = mg.Annealer(ctx, 4736, seed=3 + 1000 * rank)
coop = DistributedCooperativeAnnealer(an, dist, f"cuda:{local}", group=8, seed=1)
c0 = coop.start(None)
t0 = time.time()
for r in range(rounds):
    rec = coop.round(cycle_budget=int(ms * 1.965e6), temperature=4096.0 * (1.0 - (r + 1) / rounds))
    digest = hashlib.sha256(coop.slab().tobytes()).digest()[:8]
    t = torch.tensor(list(digest) + list(int(rec["cost"]).to_bytes(8, "little")), dtype=torch.int64, device=f"cuda:{local}")
    gathered = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(gathered, t)
    assert all(torch.equal(g, gathered[0]) for g in gathered), "ranks disagree on the slab"
    if rank == 0:
        print(rec["round"], "t=%.1fs" % (time.time() - t0), "bytes %.1f" % (rec["cost"] / 16384 + 18), rec["kept"], "regions", rec["regions"],
              "evals", rec["evals"], "run %.2fs merge %.2fs bcast %.2fs" % (rec["run_s"], rec["merge_s"], rec["broadcast_s"]), flush=True)
if rank == 0:
    stream = ctx.encode_slab(coop.slab())
    assert lzma.decompress(stream, format=lzma.FORMAT_ALONE) == data
    print("final .lzma", len(stream), "bytes on", world, "GPUs; round-trips", flush=True)
an.close()
dist.barrier()
dist.destroy_process_group()
