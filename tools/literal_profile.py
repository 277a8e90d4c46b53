"""Short bench-like run for ncu: 1 MiB mixed corpus, all-literal start, two launches with an exact packet budget
(reproducible under ncu's instrumented passes, unlike a clock budget)."""
import sys
sys.path.insert(0, '.')
import megalania_b200 as mg
from tools import corpus
n = 1 << 20
data = corpus.make("mixed", n)
ctx = mg.Context(data)
an = mg.Annealer(ctx, ctx.full_wave(), seed=5)
an.set_slab(None)
budget = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
for it in range(2):
    st = an.run(1000, first_eval=mg.CONTINUE_EVALS, suspend=True, packet_budget=budget)
    print(it, round(st["kernel_ms"], 1), st["evals"], st["packets_scored"], "finder share %.3f" % (st["finder_cycles"] / max(1, st["chain_cycles"])),
          "algorithmic_bytes", st["slab_bytes_read"] + st["checkpoint_bytes"] + 16 * st["edits"])
