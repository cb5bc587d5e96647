#!/bin/bash
# where the folded MDyEncoder spends its time: in-graph conv table + ncu launch list of one step
mkdir -p gpurun_out
UAVDET_BENCH_DEBUG=1 python bench.py --model rtm-infer --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/s27_rtm.json 2> gpurun_out/s27_rtm_table.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/s27_rtm_launches.csv \
  python bench.py --model rtm-infer --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/s27_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/s27_rtm_launches.csv 40 > gpurun_out/s27_rtm_summary.txt; head -30 gpurun_out/s27_rtm_summary.txt
