#!/bin/bash
# ncu launch lists (gpu__time_duration) of one warm step of every benchmark configuration, final round-2 binaries
mkdir -p gpurun_out
for m in baseline dysoem dyyolo rtm-infer; do
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
      --log-file gpurun_out/s10_launches_$m.csv python tools/profile_step_launches.py $m > gpurun_out/s10_ncu_$m.log 2>&1
  python tools/launch_summary.py gpurun_out/s10_launches_$m.csv 40 > gpurun_out/s10_launch_summary_$m.txt 2>&1
  echo "== $m"; head -12 gpurun_out/s10_launch_summary_$m.txt
done
