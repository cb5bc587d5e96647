#!/bin/bash
# GroupNorm folds of the MDyEncoder: new kernel / encoder tests, RTM tests, A/B of the inference bench
mkdir -p gpurun_out
python -m pytest tests/ -m gpu -x -q -k "dwdynconv or groupnorm or mdy_encoder or rtm or linear" 2>&1 | tail -15 > gpurun_out/s26_tests.log; cat gpurun_out/s26_tests.log
python bench.py --model rtm-infer > gpurun_out/s26_rtm_fold.json 2> gpurun_out/s26_rtm_fold.err
UAVDET_RTM_NO_GN_FOLD=1 python bench.py --model rtm-infer > gpurun_out/s26_rtm_nofold.json 2> gpurun_out/s26_rtm_nofold.err
python -c "
import json
for k in ('fold','nofold'):
    try:
        d=json.load(open('gpurun_out/s26_rtm_%s.json'%k)); print(k, round(d['value'],1), round(d['ms_per_step'],2))
    except Exception as e: print(k, 'failed', e)"
tail -5 gpurun_out/s26_rtm_fold.err
