#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_models.py tests/test_gpu_real_shapes.py -x -q -k "dyyolo or dysoem" 2>&1 | tail -15 > gpurun_out/r02_pytest_dyn.log
for m in rtm-infer dyyolo dysoem; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_$m.csv python tools/profile_step_launches.py $m > gpurun_out/r02_launches_$m.log 2>&1
done
tail -4 gpurun_out/r02_pytest_dyn.log
