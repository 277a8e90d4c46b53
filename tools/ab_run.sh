#!/bin/bash
# GPU parity tests, then the short bench-like run with and without the literal queues (run under gpurun)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/ab_tests.log 2>&1; tail -15 gpurun_out/ab_tests.log
python tools/variant_run.py 2>&1 | tail -2
MEGALANIA_NO_LITQ=1 python tools/variant_run.py 2>&1 | tail -2
