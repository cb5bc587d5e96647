#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/ -m gpu -x -q -k "pair or rtm or mdy" 2>&1 | tail -8 > gpurun_out/s32_tests.log; cat gpurun_out/s32_tests.log
python bench.py --model rtm-infer --no-cpu-baseline > gpurun_out/s32_rtm_pair.json 2> gpurun_out/s32_rtm_pair.err
UAVDET_RTM_NO_PAIR_CONV=1 python bench.py --model rtm-infer --no-cpu-baseline > gpurun_out/s32_rtm_nopair.json 2> gpurun_out/s32_rtm_nopair.err
UAVDET_BENCH_DEBUG=1 python bench.py --model rtm-infer --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/s32_rtm.json 2> gpurun_out/s32_rtm_table.txt
python -c "
import json
for k in ('pair','nopair'):
    try:
        d=json.load(open('gpurun_out/s32_rtm_%s.json'%k)); print(k, round(d['value'],1), round(d['ms_per_step'],2))
    except Exception as e: print(k, 'failed', e)"
tail -3 gpurun_out/s32_rtm_pair.err
