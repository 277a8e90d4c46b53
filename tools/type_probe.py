"""Packet type mix of an annealed slab: python tools/type_probe.py [n] [kind] [steps]"""
import sys
sys.path.insert(0, '.')
import numpy as np
import megalania_b200 as mg
from oracle import oracle_lib
from tools import corpus
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
kind = sys.argv[2] if len(sys.argv) > 2 else "text"
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
data = corpus.make(kind, n)
port = oracle_lib.Port()
greedy = port.greedy_slab(data)
ctx = mg.Context(data)
an = mg.Annealer(ctx, 4736, seed=3)
an.set_slab(greedy)


def mix(slab):
    cnt = {}
    p = 0
    while p < n:
        t = int(slab[p]["type"])
        cnt[t] = cnt.get(t, 0) + 1
        p += int(slab[p]["len"])
    return cnt


print("greedy", mix(greedy), "cost", port.slab_cost(data, greedy) / 16384)
for s in range(steps):
    st = an.run(1_000_000, step=2, first_eval=mg.CONTINUE_EVALS, suspend=True, cycle_budget=2_000_000_000)
    cur, best = an.costs()
    b = int(np.where(best > 0, best, np.iinfo(np.uint64).max).argmin())
    print(s, "evals", st["evals"], "best bytes", best[b] / 16384, mix(an.get_slab(b, best=True)), "accepted", st["accepted"])
