import sys
sys.path.insert(0,'.')
import megalania_b200 as mg
from tools import corpus
n=1<<20
ctx=mg.Context(corpus.make("mixed",n))
an=mg.Annealer(ctx, 4736, seed=5)
an.set_slab(None)
for it in range(2):
    st=an.run(1000, packet_budget=2000000, first_eval=mg.CONTINUE_EVALS)
    print(it, round(st['kernel_ms'],1), st['evals'], st['new_best'])
