#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/ -m gpu -x -q -k "dwdynconv or rtm or mdy" 2>&1 | tail -3
python bench.py --model rtm-infer --no-cpu-baseline > gpurun_out/s36_rtm.json 2> gpurun_out/s36_rtm.err
python -c "
import json
d=json.load(open('gpurun_out/s36_rtm.json')); print('rtm', round(d['value'],1), round(d['ms_per_step'],2))"
python tools/run_membound_kernels.py 2>/dev/null | grep -i "dwdyn" | head -12
