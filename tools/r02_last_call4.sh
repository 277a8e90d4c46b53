#!/bin/bash
# Steady-state hardware counters of the committed kernel (the last half minute of the round's GPU budget).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 30 ncu --section SpeedOfLight --section WarpStateStats --section SchedulerStats --section MemoryWorkloadAnalysis --section ComputeWorkloadAnalysis --section InstructionStats --clock-control none -k regex:anneal_kernel -s 1 -c 1 -f -o gpurun_out/p_steady python tools/steady_profile.py > gpurun_out/p_steady.log 2>&1
tail -3 gpurun_out/p_steady.log; echo "t=$SECONDS"
