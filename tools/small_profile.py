"""4 KiB exec-like input, thousands of chains (BASELINE config 3 shape), for ncu."""
import sys
sys.path.insert(0, '.')
import megalania_b200 as mg
from tools import corpus
n = 4096
data = corpus.make("binary", n)
ctx = mg.Context(data)
an = mg.Annealer(ctx, 4736, seed=5)
an.set_slab(None)
for it in range(4):
    st = an.run(1000000, first_eval=mg.CONTINUE_EVALS, suspend=True, cycle_budget=100_000_000)
    print(it, round(st["kernel_ms"], 1), st["evals"], st["packets_scored"], st["attempts"])
