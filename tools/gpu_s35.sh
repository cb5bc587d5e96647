#!/bin/bash
mkdir -p gpurun_out
for m in dysoem dyyolo; do
python bench.py --model $m --no-cpu-baseline > gpurun_out/s35_${m}_pair.json 2> gpurun_out/s35_${m}_pair.err
UAVDET_NO_PAIR_CONV=1 python bench.py --model $m --no-cpu-baseline > gpurun_out/s35_${m}_nopair.json 2> gpurun_out/s35_${m}_nopair.err
done
python -c "
import json
for m in ('dysoem','dyyolo'):
  for k in ('pair','nopair'):
    try:
        d=json.load(open('gpurun_out/s35_%s_%s.json'%(m,k))); print(m, k, round(d['value'],1), round(d['ms_per_step'],3))
    except Exception as e: print(m, k, 'failed', e)"
