#!/bin/bash
# GroupNorm fold, per-image scalar form: tests + A/B + conv table
mkdir -p gpurun_out
python -m pytest tests/ -m gpu -x -q -k "dwdynconv or groupnorm or mdy_encoder or rtm or linear or affine or gelu" 2>&1 | tail -15 > gpurun_out/s29_tests.log; cat gpurun_out/s29_tests.log
python bench.py --model rtm-infer --no-cpu-baseline > gpurun_out/s29_rtm_fold.json 2> gpurun_out/s29_rtm_fold.err
UAVDET_RTM_NO_GN_FOLD=1 python bench.py --model rtm-infer --no-cpu-baseline > gpurun_out/s29_rtm_nofold.json 2> gpurun_out/s29_rtm_nofold.err
UAVDET_BENCH_DEBUG=1 python bench.py --model rtm-infer --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/s29_rtm.json 2> gpurun_out/s29_rtm_table.txt
python -c "
import json
for k in ('fold','nofold'):
    try:
        d=json.load(open('gpurun_out/s29_rtm_%s.json'%k)); print(k, round(d['value'],1), round(d['ms_per_step'],2))
    except Exception as e: print(k, 'failed', e)"
tail -3 gpurun_out/s29_rtm_fold.err
