#!/bin/bash
# The drop-in CLI with a greedy starting slab: size after a fixed annealing budget (run under gpurun)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_gpu_config2.py -m gpu -x -q -k "greedy" 2>&1 | tail -3
python - <<'P'
from tools import corpus
open('/tmp/t64.bin','wb').write(corpus.make('text',65536))
open('/tmp/m1.bin','wb').write(corpus.make('mixed',1<<20))
P
CLI=megalania_b200/_build/megalania
for args in "--chains 3848 --time 8 --round-ms 250 /tmp/t64.bin" "--chains 3848 --time 8 --round-ms 250 --greedy 64 /tmp/t64.bin" "--chains 3848 --time 8 --round-ms 250 --greedy 1024 /tmp/m1.bin" "--chains 3848 --time 2 --round-ms 250 --greedy 1024 /tmp/m1.bin"; do
  f=${args##* }
  s=$(date +%s.%N); $CLI $args > /tmp/out.lzma 2> /tmp/err.txt; e=$(date +%s.%N); tail -2 /tmp/err.txt; echo "wall $(echo "$e - $s" | bc) s"
  echo "$args -> $(stat -c %s /tmp/out.lzma) bytes; round trip: $(xz --format=lzma -dc /tmp/out.lzma | cmp - $f && echo ok)"
done
