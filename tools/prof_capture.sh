set -x
python tools/literal_profile.py > gpurun_out/p0_plain.log 2>&1 || exit 1
cat gpurun_out/p0_plain.log
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:anneal_kernel -s 1 -c 1 -f -o gpurun_out/p0_anneal python tools/literal_profile.py > gpurun_out/p0_ncu.log 2>&1
tail -3 gpurun_out/p0_ncu.log
python tools/split_probe.py 1048576 mixed 3996 0 suspend 786000000 > gpurun_out/p0_split.log 2>&1; cat gpurun_out/p0_split.log
