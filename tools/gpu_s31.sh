#!/bin/bash
# GroupNorm fold v3 (per-warp additive vector, pairs staged in shared memory): microbench, all GPU tests, RTM A/B, smoke
mkdir -p gpurun_out
python tools/bench_gn_fold_conv.py 2>&1 | tail -2
python -m pytest tests/ -m gpu -x -q 2>&1 | tail -3 > gpurun_out/s31_tests.log; cat gpurun_out/s31_tests.log
python bench.py --model rtm-infer --no-cpu-baseline > gpurun_out/s31_rtm_fold.json 2> gpurun_out/s31_rtm_fold.err
UAVDET_RTM_NO_GN_FOLD=1 python bench.py --model rtm-infer --no-cpu-baseline > gpurun_out/s31_rtm_nofold.json 2> gpurun_out/s31_rtm_nofold.err
python -c "
import json
for k in ('fold','nofold'):
    try:
        d=json.load(open('gpurun_out/s31_rtm_%s.json'%k)); print(k, round(d['value'],1), round(d['ms_per_step'],2))
    except Exception as e: print(k, 'failed', e)"
