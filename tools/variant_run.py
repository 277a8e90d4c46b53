"""One timed configuration for tools/variants.sh: 1 MiB mixed corpus, all-literal start, a full wave of chains,
four 400 ms launches (the first is warm-up); prints successful evaluations per second of device time."""
import sys
sys.path.insert(0, '.')
import megalania_b200 as mg
from tools import corpus
n = 1 << 20
data = corpus.make("mixed", n)
ctx = mg.Context(data)
an = mg.Annealer(ctx, ctx.full_wave(), seed=5)
an.set_slab(None)
ev = ms = bits = 0
for it in range(4):
    st = an.run(1000, first_eval=mg.CONTINUE_EVALS, suspend=True, cycle_budget=int(0.4 * 1.965e9))
    if it:
        ev += st["evals"]; ms += st["kernel_ms"]; bits += st["bits_scored"]
print(f"chains {an.chains}  {ev / (ms / 1e3):.0f} evals/s  {bits / (ms / 1e3) / 3.10208e12:.4f} of the bank roofline  finder share {st['finder_cycles'] / max(1, st['chain_cycles']):.3f}")
