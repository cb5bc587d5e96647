#!/bin/bash
# pixel-pair GEMM for the thin 32 -> 64 training layer: tests, A/B of the default bench with the in-graph conv table
mkdir -p gpurun_out
python -m pytest tests/ -m gpu -x -q 2>&1 | tail -4 > gpurun_out/s34_tests.log; cat gpurun_out/s34_tests.log
UAVDET_BENCH_DEBUG=1 python bench.py --no-cpu-baseline > gpurun_out/s34_pair.json 2> gpurun_out/s34_pair_table.txt
UAVDET_NO_PAIR_CONV=1 UAVDET_BENCH_DEBUG=1 python bench.py --no-cpu-baseline > gpurun_out/s34_nopair.json 2> gpurun_out/s34_nopair_table.txt
python -c "
import json
for k in ('pair','nopair'):
    try:
        d=json.load(open('gpurun_out/s34_%s.json'%k)); print(k, round(d['value'],1), round(d['ms_per_step'],3), 'frac', round(d['roofline']['frac'],3))
    except Exception as e: print(k, 'failed', e)"
grep convtimer gpurun_out/s34_pair_table.txt | grep igemm | tail -12
grep convtimer gpurun_out/s34_nopair_table.txt | grep igemm | tail -12
