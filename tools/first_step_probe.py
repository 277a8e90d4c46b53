import sys, time
sys.path.insert(0,'.')
import numpy as np
import megalania_b200 as mg
from tools import corpus
n=1<<20; chains=4736
data=corpus.make("mixed",n)
ctx=mg.Context(data)
an=mg.Annealer(ctx, chains, seed=5)
for rnd in range(2):
    an.set_slab(None)
    for it in range(3):
        st=an.run(1000, packet_budget=4000000, first_eval=mg.CONTINUE_EVALS)
        print(rnd, it, "kernel_ms",round(st['kernel_ms'],1),"evals",st['evals'],"att",st['attempts'],"acc",st['accepted'],"newbest",st['new_best'],"bits/s %.3g"%(st['bits_scored']/(st['kernel_ms']/1e3)),"pk/att",st['packets_scored']//st['attempts'], "cand", st['finder_candidates']//max(1,st['finder_calls']), "rejoined", st["rejoined"], "slabB", st["slab_bytes_read"])
# clock warm-up hypothesis: keep the GPU busy with an unrelated annealer right before step 0
an2=mg.Annealer(ctx, 4736, seed=9, track_best=False); an2.set_slab(None)
an.set_slab(None)
an2.run(1000, packet_budget=4000000); an2.run(1000, packet_budget=4000000)
st=an.run(1000, packet_budget=4000000, first_eval=mg.CONTINUE_EVALS)
print("after warm GPU: step0 kernel_ms", round(st['kernel_ms'],1), "evals", st['evals'], "newbest", st['new_best'])
st=an.run(1000, packet_budget=4000000, first_eval=mg.CONTINUE_EVALS)
print("step1 kernel_ms", round(st['kernel_ms'],1))
