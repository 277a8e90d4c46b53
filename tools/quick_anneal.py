import sys, time
sys.path.insert(0,'.')
import numpy as np
import megalania_b200 as mg
from tools import corpus
n=int(sys.argv[1]); chains=int(sys.argv[2]); evals=int(sys.argv[3]); kind=sys.argv[4] if len(sys.argv)>4 else "mixed"
budget=int(sys.argv[5]) if len(sys.argv)>5 else 0
data=corpus.make(kind,n)
t=time.time(); ctx=mg.Context(data); print("ctx",time.time()-t)
print("chain bytes", ctx.chain_bytes(chains=chains, track_best=1))
t=time.time(); an=mg.Annealer(ctx, chains, seed=5); an.set_slab(None); print("create+set",time.time()-t)
for it in range(3):
    t=time.time(); st=an.run(evals, packet_budget=budget, first_eval=mg.CONTINUE_EVALS); dt=time.time()-t
    print(it, "wall",round(dt,3),"kernel_ms",round(st['kernel_ms'],1),"evals/s",round(st['evals']/(st['kernel_ms']/1e3),1),
      "bits/s %.3g"%(st['bits_scored']/(st['kernel_ms']/1e3)), "pk/eval",st['packets_scored']//max(1,st['attempts']), "att",st['attempts'],"acc",st['accepted'],"cand/find",st['finder_candidates']//max(1,st['finder_calls']), "finds", st["finder_calls"], "rejoined", st["rejoined"])
cur,best=an.costs(); print("best", best.min(), "cur mean", cur.mean())
