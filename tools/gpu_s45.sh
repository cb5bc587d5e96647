#!/bin/bash
# RTMUAVDet launch evidence from the very last binaries
mkdir -p gpurun_out
UAVDET_BENCH_DEBUG=1 python bench.py --model rtm-infer --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/s45_dbg_rtm.json 2> gpurun_out/r02_final_conv_launch_table_rtm-infer.txt
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv \
   --log-file gpurun_out/r02_launches_rtm-infer_step.csv python tools/profile_step_launches.py rtm-infer > gpurun_out/s45_ncu.log 2>&1
python tools/launch_summary_dram.py gpurun_out/r02_launches_rtm-infer_step.csv 28 > gpurun_out/r02_launch_summary_rtm-infer_step.txt
head -12 gpurun_out/r02_launch_summary_rtm-infer_step.txt
