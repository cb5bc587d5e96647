#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r02_pytest_c11.log
UAVDET_BENCH_DEBUG=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_c11_bench.json 2> gpurun_out/r02_c11_table.txt
REPS=2 python tools/bench_wgrad.py wgrad 2 > gpurun_out/r02_c11_wgrad_plain.log 2>&1 && \
REPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:wgrad_kernel -s 2 -c 1 -f -o gpurun_out/r02_wgrad20 python tools/bench_wgrad.py wgrad 2 > gpurun_out/r02_c11_ncu.log 2>&1
tail -3 gpurun_out/r02_pytest_c11.log
