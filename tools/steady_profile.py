"""Short bench-like run for ncu's hardware counters in the steady state: 1 MiB mixed corpus, all-literal start,
two clock-boxed launches (all warps busy until the end; only rates are meaningful, counts differ per pass)."""
import sys
sys.path.insert(0, '.')
import megalania_b200 as mg
from tools import corpus
n = 1 << 20
data = corpus.make("mixed", n)
ctx = mg.Context(data)
an = mg.Annealer(ctx, ctx.full_wave(), seed=5)
an.set_slab(None)
for it in range(2):
    st = an.run(1000, first_eval=mg.CONTINUE_EVALS, suspend=True, cycle_budget=300_000_000)
    print(it, round(st["kernel_ms"], 1), st["evals"], st["packets_scored"])
