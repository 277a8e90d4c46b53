"""Short run from a greedy-parsed (match-heavy) slab for ncu: python tools/mature_profile.py [n] [kind]"""
import sys
sys.path.insert(0, '.')
import megalania_b200 as mg
from oracle import oracle_lib
from tools import corpus
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
kind = sys.argv[2] if len(sys.argv) > 2 else "text"
data = corpus.make(kind, n)
greedy = oracle_lib.Port().greedy_slab(data)
ctx = mg.Context(data)
an = mg.Annealer(ctx, 4736, seed=11)
an.set_slab(greedy)
for it in range(2):
    st = an.run(100000, first_eval=mg.CONTINUE_EVALS, step=2, suspend=True, cycle_budget=150_000_000)
    print(it, round(st["kernel_ms"], 1), st["evals"], st["packets_scored"])
