import sys
sys.path.insert(0, '.')
import megalania_b200 as mg
from oracle import oracle_lib
from tools import corpus
n = 65536
data = corpus.make("text", n)
greedy = oracle_lib.Port().greedy_slab(data)
ctx = mg.Context(data)
an = mg.Annealer(ctx, 4736, seed=11)
an.set_slab(greedy)
for it in range(2):
    st = an.run(1000, packet_budget=100_000, first_eval=mg.CONTINUE_EVALS, step=2)
    print(it, round(st["kernel_ms"], 1), st["evals"])
