"""Second timed configuration for kernel A/B runs: a match-heavy slab (64 KiB text, the device's greedy parse as
every chain's starting slab), a full wave of chains, three 150 ms launches (the first is warm-up); prints successful
evaluations per second of device time."""
import sys
sys.path.insert(0, '.')
import megalania_b200 as mg
from tools import corpus
n = 65536
data = corpus.make("text", n)
ctx = mg.Context(data)
an = mg.Annealer(ctx, ctx.full_wave(), seed=11)
an.set_slab(None)
an.greedy_init(16, 0)
an.broadcast_chain(0)
ev = ms = 0
for it in range(3):
    st = an.run(100000, first_eval=mg.CONTINUE_EVALS, step=2, suspend=True, cycle_budget=int(0.15 * 1.965e9))
    if it:
        ev += st["evals"]; ms += st["kernel_ms"]
print(f"mature chains {an.chains}  {ev / (ms / 1e3):.0f} evals/s  finder share {st['finder_cycles'] / max(1, st['chain_cycles']):.3f}")
