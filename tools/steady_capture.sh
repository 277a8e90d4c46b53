cd /root/repo
timeout 900 ncu --section SpeedOfLight --section WarpStateStats --section SchedulerStats --section MemoryWorkloadAnalysis --section ComputeWorkloadAnalysis --section InstructionStats --clock-control none -k regex:anneal_kernel -s 1 -c 1 -f -o gpurun_out/steady python tools/steady_profile.py > gpurun_out/steady.log 2>&1
tail -3 gpurun_out/steady.log
