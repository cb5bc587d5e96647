#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/ -m gpu -x -q -k "rtm or mdy or pair" 2>&1 | tail -3
python bench.py --model rtm-infer --no-cpu-baseline > gpurun_out/s40_rtm.json 2> gpurun_out/s40_rtm.err
UAVDET_RTM_NO_PAIR_CONV=1 python bench.py --model rtm-infer --no-cpu-baseline > gpurun_out/s40_rtm_nopair.json 2> gpurun_out/s40_rtm_nopair.err
python -c "
import json
for k in ('','_nopair'):
    d=json.load(open('gpurun_out/s40_rtm%s.json'%k)); print('rtm'+k, round(d['value'],1), round(d['ms_per_step'],2))"
tail -3 gpurun_out/s40_rtm.err
