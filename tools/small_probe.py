import sys
sys.path.insert(0,'.')
import megalania_b200 as mg
from tools import corpus
n=4096
ctx=mg.Context(corpus.make("binary",n))
an=mg.Annealer(ctx, 4736, seed=5)
an.set_slab(None)
for it in range(3):
    st=an.run(100)
    print(it, round(st['kernel_ms'],1), st['evals'], st['attempts'])
