#!/bin/bash
# What is left of the round's GPU budget (2.5 minutes): 26 warps x 72 registers against 24 warps x 80 registers now
# that the window loop no longer waits on HBM, the steady-state hardware counters of the faster build, and - only if
# the 24-warp build wins - the trace-parity tests on it.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
V=megalania_b200/_build/variants
rm -f $O/m_ab.log
for v in v0 w24 v0 w24; do
  echo "== variant $v" >> $O/m_ab.log
  MEGALANIA_CUDA_LIB=$V/lib_$v.so timeout 30 python tools/variant_run.py >> $O/m_ab.log 2>&1
done
cat $O/m_ab.log
W=$(python - <<'P'
import re, collections
rate, cur = collections.defaultdict(list), None
for line in open("gpurun_out/m_ab.log"):
    m = re.match(r"== variant (\w+)", line)
    if m: cur = m.group(1)
    m = re.search(r"(\d+) evals/s", line)
    if m and cur: rate[cur].append(int(m.group(1)))
mean = {k: sum(v) / len(v) for k, v in rate.items()}
print("w24" if mean.get("w24", 0) > 1.015 * mean.get("v0", 1e18) else "v0")
P
)
echo "winner $W" | tee $O/m_winner.txt
export MEGALANIA_CUDA_LIB=$PWD/$V/lib_$W.so
echo "t=$SECONDS after A/B"
timeout 60 ncu --section SpeedOfLight --section WarpStateStats --section SchedulerStats --section MemoryWorkloadAnalysis --section ComputeWorkloadAnalysis --section InstructionStats --clock-control none -k regex:anneal_kernel -s 1 -c 1 -f -o $O/m_steady python tools/steady_profile.py > $O/m_steady.log 2>&1
tail -3 $O/m_steady.log
echo "t=$SECONDS after ncu"
if [ "$W" = "w24" ] && [ $SECONDS -lt 60 ]; then
  timeout 75 python -u -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_config2.py -m gpu -x -v \
    -k "hello or cost_model or anneal_trace or suspended or clock_boxed_steps or 1mib_cost or long_run or edge_windows or early_exit or 1024_chains" \
    > $O/m_tests.log 2>&1
  echo "pytest rc=$?" >> $O/m_tests.log; tail -4 $O/m_tests.log
fi
echo "t=$SECONDS done"
