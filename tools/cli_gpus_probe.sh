#!/bin/bash
# The C host with one thread per GPU (--gpus N): cooperative rounds over the library's NCCL collectives.
#   gpurun --gpus 2 -- bash tools/cli_gpus_probe.sh 2
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
python - <<'P'
from tools import corpus
open('/tmp/m1.bin','wb').write(corpus.make('mixed',1<<20))
P
timeout 120 megalania_b200/_build/megalania --gpus $N --chains 1924 --time 3 --round-ms 500 --greedy 1024 /tmp/m1.bin > /tmp/m1.lzma 2> gpurun_out/cli_gpus.err
rc=$?
echo "exit $rc bytes $(stat -c %s /tmp/m1.lzma) round trip: $(xz --format=lzma -dc /tmp/m1.lzma 2>&1 | cmp - /tmp/m1.bin 2>&1 && echo ok)" | tee gpurun_out/cli_gpus.out
tail -2 gpurun_out/cli_gpus.err
