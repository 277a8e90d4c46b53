"""Where the end-to-end (host buffers in/out) time goes: python tools/e2e_breakdown.py [n] [chains]"""
import sys, time
sys.path.insert(0, '.')
import numpy as np
import megalania_b200 as mg
from tools import corpus
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
chains = int(sys.argv[2]) if len(sys.argv) > 2 else 4736
data = corpus.make("mixed", n)
mg.load_library()
for rep in range(2):
    t = [time.perf_counter()]
    ctx = mg.Context(data); t.append(time.perf_counter())
    an = mg.Annealer(ctx, chains, seed=1); t.append(time.perf_counter())
    an.set_slab(None); t.append(time.perf_counter())
    st = an.run(1000, packet_budget=4_000_000); t.append(time.perf_counter())
    cur, best = an.costs(); slab = an.get_slab(int(best.argmin()), best=True); t.append(time.perf_counter())
    an.close(); ctx.close(); t.append(time.perf_counter())
    names = ["ctx_create", "anneal_create", "set_slab", "run", "read_best", "destroy"]
    print(rep, {k: round(1e3 * (b - a), 1) for k, a, b in zip(names, t, t[1:])}, "evals", st["evals"], "kernel_ms", round(st["kernel_ms"], 1))
