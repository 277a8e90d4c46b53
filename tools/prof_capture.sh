#!/bin/bash
# ncu --set full capture of the second anneal_kernel launch of tools/literal_profile.py (run under gpurun)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
tag=${1:-p}
python tools/literal_profile.py > gpurun_out/${tag}_plain.log 2>&1 || { cat gpurun_out/${tag}_plain.log; exit 1; }
cat gpurun_out/${tag}_plain.log
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:anneal_kernel -s 1 -c 1 -f -o gpurun_out/${tag}_anneal python tools/literal_profile.py > gpurun_out/${tag}_ncu.log 2>&1
tail -3 gpurun_out/${tag}_ncu.log
