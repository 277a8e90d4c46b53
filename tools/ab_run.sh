#!/bin/bash
# GPU parity tests, then the short bench-like run (run under gpurun); "notests" skips the tests
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
if [ "$1" != "notests" ]; then
  python -m pytest tests -m gpu -x -q > gpurun_out/ab_tests.log 2>&1; tail -15 gpurun_out/ab_tests.log
fi
python tools/variant_run.py 2>&1 | tail -2
