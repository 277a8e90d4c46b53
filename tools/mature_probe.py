"""Early-exit effect on a mature (greedy-parsed) slab: python tools/mature_probe.py [n] [kind]"""
import sys, time
sys.path.insert(0, '.')
import numpy as np
import megalania_b200 as mg
from oracle import oracle_lib
from tools import corpus
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
kind = sys.argv[2] if len(sys.argv) > 2 else "text"
data = corpus.make(kind, n)
port = oracle_lib.Port()
t = time.time(); greedy = port.greedy_slab(data); print("greedy", round(time.time() - t, 1), "s, live", port.live_count(greedy))
ctx = mg.Context(data)
for early in (False, True):
    an = mg.Annealer(ctx, 4736, seed=11)
    an.set_slab(greedy)
    for it in range(3):
        st = an.run(1000, packet_budget=2_000_000 if n >= 1 << 20 else 400_000, first_eval=mg.CONTINUE_EVALS, early_exit=early, step=2)
        print("early" if early else "full ", it, "kernel_ms", round(st["kernel_ms"], 1), "evals", st["evals"], "evals/s", round(st["evals"] / st["kernel_ms"] * 1e3),
              "rejoined", st["rejoined"], "pk/att", st["packets_scored"] // st["attempts"], "acc", st["accepted"])
    cur, best = an.costs(); print("  best", best.min(), "mean cur", cur.mean())
    an.close()
