"""Compression progress of the cooperative search: python tools/coop_probe.py [n] [kind] [rounds] [ms] [group] [start]"""
import lzma
import subprocess
import sys
import time
sys.path.insert(0, '.')
import numpy as np
import megalania_b200 as mg
from megalania_b200.cooperative import CooperativeAnnealer
from tools import corpus
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
kind = sys.argv[2] if len(sys.argv) > 2 else "text"
rounds = int(sys.argv[3]) if len(sys.argv) > 3 else 10
ms = float(sys.argv[4]) if len(sys.argv) > 4 else 500.0
group = int(sys.argv[5]) if len(sys.argv) > 5 else 8
start = sys.argv[6] if len(sys.argv) > 6 else "literal"
temp0 = float(sys.argv[7]) if len(sys.argv) > 7 else 0.0  # initial temperature (1/2048 bit), cooled linearly to 0
ladder = len(sys.argv) > 8 and sys.argv[8] == "ladder"
data = corpus.make(kind, n)
xz = len(subprocess.run(["xz", "-9e", "--format=lzma", "-c"], input=data, stdout=subprocess.PIPE).stdout)
print("xz -9e --format=lzma:", xz, "bytes")
ctx = mg.Context(data)
an = mg.Annealer(ctx, 4736, seed=3)
coop = CooperativeAnnealer(an, group=group, seed=1)
init = None
if start == "greedy":
    from oracle import oracle_lib
    init = oracle_lib.Port().greedy_slab(data)
c0 = coop.start(init)
print("start", start, c0 / 16384 + 18)
t0 = time.time()
for r in range(rounds):
    rec = coop.round(cycle_budget=int(ms * 1.965e6), temperature=temp0 * (1.0 - (r + 1) / rounds), ladder=ladder)
    print(rec["round"], "t=%.1fs" % (time.time() - t0), "bytes %.1f" % (rec["cost"] / 16384 + 18), "merged %.1f single %.1f" % (rec["merged"] / 16384 + 18, rec["best_single"] / 16384 + 18),
          rec["kept"], "regions", rec["regions"], "evals", rec["evals"],
          "run %.2fs merge %.2fs bcast %.2fs" % (rec["run_s"], rec["merge_s"], rec["broadcast_s"]))
stream = ctx.encode_slab(coop.slab())
assert lzma.decompress(stream, format=lzma.FORMAT_ALONE) == data
slab = coop.slab()
mix, p_ = {}, 0
while p_ < n:
    t_ = int(slab[p_]["type"]); mix[t_] = mix.get(t_, 0) + 1; p_ += int(slab[p_]["len"])
print("packet mix {LITERAL 1, MATCH 2, SHORT_REP 3, LONG_REP 4}:", mix)
print("final .lzma", len(stream), "bytes; xz -9e", xz, "; round-trips")
an.close()
