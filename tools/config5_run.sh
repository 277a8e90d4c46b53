#!/bin/bash
# BASELINE configs[4]: 16 MiB synthetic corpus, large dictionary window, fixed time budget on N GPUs through the
# drop-in C CLI; size beside the reference harness on one host core at the same wall clock and xz -9e.
#   gpurun --gpus 8 -- bash tools/config5_run.sh 8 30
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-8}; T=${2:-30}; CH=${3:-256}
python - <<'P'
from tools import corpus
open('/tmp/c16.bin','wb').write(corpus.make('corpus16', 16 << 20))
P
CLI=megalania_b200/_build/megalania
s=$(date +%s)
$CLI --gpus $N --chains $CH --time $T --round-ms 1000 --greedy 4096 --max-occ 512 --window 16777216 /tmp/c16.bin > /tmp/c16.lzma 2> gpurun_out/config5_cli.err
rc=$?
e=$(date +%s)
tail -3 gpurun_out/config5_cli.err
python - "$N" "$T" "$CH" "$((e - s))" "$rc" <<'P'
import sys, json, lzma, time, subprocess
sys.path.insert(0, '.')
n_gpus, budget, chains, wall, rc = (int(x) for x in sys.argv[1:6])
data = open('/tmp/c16.bin', 'rb').read()
out = open('/tmp/c16.lzma', 'rb').read()
rec = {"config": "BASELINE configs[4]: 16 MiB synthetic corpus (tools/corpus.py corpus16, seed 99)", "n_gpus": n_gpus, "chains_per_gpu": chains,
       "command": f"megalania --gpus {n_gpus} --chains {chains} --time {budget} --round-ms 1000 --greedy 4096 --max-occ 512 --window 16777216 <file>",
       "exit_code": rc, "wall_s": wall, "bytes": len(out)}
try:
    rec["round_trip"] = lzma.decompress(out, format=lzma.FORMAT_ALONE) == data
    rec["header_dict_size"] = int.from_bytes(out[1:5], "little")
except Exception as ex:
    rec["round_trip"] = False
    rec["error"] = str(ex)
t0 = time.time()
rec["xz_9e_bytes"] = len(lzma.compress(data, format=lzma.FORMAT_ALONE, preset=9 | lzma.PRESET_EXTREME))
rec["xz_9e_s"] = round(time.time() - t0, 1)
# the reference's own loop on one host core for the same wall clock, from the all-literal slab (src/main.c:78-102)
from oracle import oracle_lib as ol
ol.build()
import os
is_ref = os.path.exists(ol.REF_SO)
lib = ol.Ref() if is_ref else ol.Port()
kw = {} if is_ref else {"rng_mode": 0}
slab = ol.literal_slab(len(data)); best = slab.copy()
t0 = time.time(); evals = 0; bc = cc = 0; first = True
while time.time() - t0 < wall:
    r = lib.anneal_epoch(data, slab, best, bc, cc, reseed=1 if first else 0, seed=1673551, evals=1, **kw)
    bc, cc = r[1], r[2]; evals += 1; first = False
stream = lib.encode_slab(data, best if bc else slab)
rec["reference_same_wall"] = {"bytes": len(stream), "evals": evals, "wall_s": round(time.time() - t0, 1), "cores": 1}
rec["no_larger_than_reference"] = rec["bytes"] <= len(stream)
rec["smaller_than_xz_9e"] = rec["bytes"] < rec["xz_9e_bytes"]
print(json.dumps(rec))
open('gpurun_out/config5.json', 'w').write(json.dumps(rec, indent=1) + "\n")
P
