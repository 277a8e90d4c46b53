"""16 MiB smoke (BASELINE config 5 shape): context, scoring, sampled top-k, a time-boxed step, one merge.
python tools/big_probe.py [n] [chains] [ms]"""
import lzma
import sys
import time
sys.path.insert(0, '.')
import numpy as np
import megalania_b200 as mg
from megalania_b200.cooperative import CooperativeAnnealer
from tools import corpus
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16 << 20
chains = int(sys.argv[2]) if len(sys.argv) > 2 else 296
ms = float(sys.argv[3]) if len(sys.argv) > 3 else 3000.0
t = time.time(); data = corpus.make("corpus16", n); print("corpus", round(time.time() - t, 1), "s")
t = time.time(); ctx = mg.Context(data); print("context + index", round(time.time() - t, 2), "s")
lit = mg.literal_slab(n)
t = time.time(); c = ctx.score_slab(lit); print("all-literal cost", c, "=", c / 16384 + 18, "bytes", round(time.time() - t, 2), "s")
pos = np.random.default_rng(1).integers(1, n - 1, 64).astype(np.uint64)
t = time.time(); pops, prices, counts = ctx.find_topk(lit, pos, state_mode=0); print("top-k at 64 positions", round(time.time() - t, 2), "s", "counts", counts[:8])
print("chain bytes", ctx.chain_bytes(chains=chains, track_best=1))
an = mg.Annealer(ctx, chains, seed=3)
coop = CooperativeAnnealer(an, group=1, seed=1)
c0 = coop.start(None)
for r in range(2):
    rec = coop.round(cycle_budget=int(ms * 1.965e6))
    print(rec["round"], "bytes %.1f" % (rec["cost"] / 16384 + 18), rec["kept"], "regions", rec["regions"], "evals", rec["evals"],
          "run %.2fs merge %.2fs bcast %.2fs" % (rec["run_s"], rec["merge_s"], rec["broadcast_s"]))
t = time.time(); stream = ctx.encode_slab_buffer(coop.slab()); print("range coder", round(time.time() - t, 2), "s", len(stream), "bytes, dict", int.from_bytes(stream[1:5], "little"))
assert lzma.decompress(stream, format=lzma.FORMAT_ALONE) == data
print("round-trips")
an.close()
