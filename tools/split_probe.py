"""Where a step's time goes: python tools/split_probe.py [n] [kind] [chains] [budget] [mature]
Prints, per step: evals/s, share of warp-busy clocks spent in the match finder, and how evenly the
chains ended (busy clocks / (chains x longest chain))."""
import sys
sys.path.insert(0, '.')
import megalania_b200 as mg
from tools import corpus
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
kind = sys.argv[2] if len(sys.argv) > 2 else "mixed"
chains = int(sys.argv[3]) if len(sys.argv) > 3 else 4736
budget = int(sys.argv[4]) if len(sys.argv) > 4 else 4_000_000
mature = len(sys.argv) > 5 and "mature" in sys.argv[5]
suspend = len(sys.argv) > 5 and "suspend" in sys.argv[5]
cycles = int(sys.argv[6]) if len(sys.argv) > 6 else 0
data = corpus.make(kind, n)
ctx = mg.Context(data)
an = mg.Annealer(ctx, chains, seed=5)
if mature:
    from oracle import oracle_lib
    an.set_slab(oracle_lib.Port().greedy_slab(data))
else:
    an.set_slab(None)
for it in range(3):
    st = an.run(1000, packet_budget=budget, first_eval=mg.CONTINUE_EVALS, suspend=suspend, cycle_budget=cycles)
    ms = st["kernel_ms"]
    print(it, "kernel_ms", round(ms, 1), "evals/s", round(st["evals"] / ms * 1e3), "att", st["attempts"],
          "pk/att", st["packets_scored"] // max(1, st["attempts"]), "cand/find", st["finder_candidates"] // max(1, st["finder_calls"]),
          "find share %.3f" % (st["finder_cycles"] / max(1, st["chain_cycles"])),
          "cyc/find", st["finder_cycles"] // max(1, st["finder_calls"]), "cyc/chunk", st["finder_cycles"] // max(1, st["finder_chunks"]),
          "cyc/packet(walk) %.1f" % ((st["chain_cycles"] - st["finder_cycles"]) / max(1, st["packets_scored"])),
          "evenness %.3f" % (st["chain_cycles"] / (chains * max(1, st["max_chain_cycles"]))),
          "kernel cycles", st["max_chain_cycles"])
