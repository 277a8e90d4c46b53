"""Match finder alone, for ncu: top-k at the first 16 Ki positions of the 1 MiB corpus + 2048 sampled ones."""
import sys
sys.path.insert(0, '.')
import numpy as np
import megalania_b200 as mg
from tools import corpus
n = 1 << 20
data = corpus.make("mixed", n)
ctx = mg.Context(data)
pos = np.concatenate([np.arange(16384, dtype=np.uint64), np.random.default_rng(7).integers(0, n, 2048).astype(np.uint64)])
lit = mg.literal_slab(n)
for it in range(2):
    ctx.find_topk(lit, pos, state_mode=0)
    print(it, ctx.find_topk_stats())
