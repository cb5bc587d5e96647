#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/ -m gpu -x -q -k "gelu or affine or rtm or mdy or groupnorm or conv_fwd" 2>&1 | tail -3
python tools/bench_gn_fold_conv.py 2>&1 | tail -2
python bench.py --model rtm-infer --no-cpu-baseline > gpurun_out/s38_rtm.json 2> gpurun_out/s38_rtm.err
python -c "
import json
d=json.load(open('gpurun_out/s38_rtm.json')); print('rtm', round(d['value'],1), round(d['ms_per_step'],2))"
