#!/bin/bash
# The last 90 seconds of the round's GPU budget: does the finder run faster when the proposal's state is parked in
# local memory around the call (variants 16, 48)?  Both timed configurations (literal-heavy 1 MiB, match-heavy 64 KiB);
# a variant must win the first by 1.5 % and not lose the second by 2 %; the trace-parity tests then run on the winner.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
V=megalania_b200/_build/variants
rm -f $O/n_ab.log
for v in v0b v16 v48; do
  echo "== variant $v" >> $O/n_ab.log
  MEGALANIA_CUDA_LIB=$V/lib_$v.so timeout 20 python tools/variant_run.py >> $O/n_ab.log 2>&1
  MEGALANIA_CUDA_LIB=$V/lib_$v.so timeout 20 python tools/variant_run_mature.py >> $O/n_ab.log 2>&1
done
cat $O/n_ab.log
W=$(python - <<'P'
import re
lit, mat, cur = {}, {}, None
for line in open("gpurun_out/n_ab.log"):
    m = re.match(r"== variant (\w+)", line)
    if m: cur = m.group(1)
    m = re.search(r"(\d+) evals/s", line)
    if m and cur: (mat if line.startswith("mature") else lit)[cur] = int(m.group(1))
best = "v0b"
for v in ("v16", "v48"):
    if v in lit and v in mat and "v0b" in lit and "v0b" in mat:
        if lit[v] > 1.015 * lit["v0b"] and mat[v] > 0.98 * mat["v0b"] and lit[v] > lit.get(best, 0): best = v
print(best)
P
)
echo "winner $W" | tee $O/n_winner.txt
echo "t=$SECONDS after A/B"
if [ "$W" != "v0b" ] && [ $SECONDS -lt 32 ]; then
  export MEGALANIA_CUDA_LIB=$PWD/$V/lib_$W.so
  timeout 52 python -u -m pytest tests/test_gpu_parity.py tests/test_gpu_config2.py tests/test_gpu_fullsize.py -m gpu -x -v \
    -k "hello or cost_model or anneal_trace or suspended or clock_boxed_steps or gives_up or long_run or edge_windows or early_exit" \
    > $O/n_tests.log 2>&1
  echo "pytest rc=$?" >> $O/n_tests.log; tail -4 $O/n_tests.log
fi
echo "t=$SECONDS done"
