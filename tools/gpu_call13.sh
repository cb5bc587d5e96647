#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r02_pytest_c13.log
UAVDET_IGEMM_RES_PREFETCH=0 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_c13_bench_nopf.json 2> gpurun_out/r02_c13_bench_nopf.err
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_c13_bench_pf.json 2> gpurun_out/r02_c13_bench_pf.err
UAVDET_IGEMM_RES_PREFETCH=0 python tools/bench_wgrad.py dgradres 6,7,9 > gpurun_out/r02_c13_dgradres_nopf.txt 2>&1
python tools/bench_wgrad.py dgradres 6,7,9 > gpurun_out/r02_c13_dgradres_pf.txt 2>&1
for m in dysoem rtm-infer; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches2_$m.csv python tools/profile_step_launches.py $m > gpurun_out/r02_launches2_$m.log 2>&1
done
tail -3 gpurun_out/r02_pytest_c13.log
