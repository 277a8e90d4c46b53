#!/bin/bash
# Round-end evidence in one gpurun call: GPU parity tests, ncu --set full of one anneal_kernel launch, a short plain
# bench and the ncu launch list of the same command.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/final_tests.log 2>&1; tail -4 gpurun_out/final_tests.log
python tools/literal_profile.py > gpurun_out/final_plain.log 2>&1; cat gpurun_out/final_plain.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:anneal_kernel -s 1 -c 1 -f -o gpurun_out/final_anneal python tools/literal_profile.py > gpurun_out/final_ncu.log 2>&1; tail -2 gpurun_out/final_ncu.log
SHORT="bench.py --steps 2 --warmup 3 --no-size --no-e2e --no-cpu --no-finder"
python $SHORT > gpurun_out/final_short_bench.json 2> gpurun_out/final_short_bench.err; tail -c 400 gpurun_out/final_short_bench.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches.csv python $SHORT > gpurun_out/final_launches.log 2>&1; tail -2 gpurun_out/final_launches.log
