#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/ -m gpu -x -q 2>&1 | tail -6 > gpurun_out/s13_tests.log
cat gpurun_out/s13_tests.log
python tools/bench_stem1x1.py > gpurun_out/s13_stem1x1.txt 2>&1; cat gpurun_out/s13_stem1x1.txt
UAVDET_BENCH_DEBUG=1 python bench.py --model rtm-infer --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/s13_rtm.json 2> gpurun_out/s13_rtm_table.txt
python bench.py --model dysoem --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/s13_dysoem.json 2>/dev/null
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/s13_baseline.json 2>/dev/null
for f in gpurun_out/s13_*.json; do python -c "
import json; d=json.load(open('$f')); print('$f', round(d['value'],1), round(d['ms_per_step'],3))"; done
