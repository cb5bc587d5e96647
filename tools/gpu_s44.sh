#!/bin/bash
# packed fp32x2 math in every AFFINE epilogue (scale / shift FMA, SiLU, residual add): microbench, RTMUAVDet, BaselineModel
mkdir -p gpurun_out
python -m pytest tests/ -m gpu -x -q -k "affine or silu or conv_fwd or conv_dgrad or rtm or head" 2>&1 | tail -3
python tools/bench_gn_fold_conv.py 2>&1 | tail -2
python bench.py --model rtm-infer --no-cpu-baseline > gpurun_out/s44_rtm.json 2> gpurun_out/s44_rtm.err
python bench.py --no-cpu-baseline > gpurun_out/s44_base.json 2> gpurun_out/s44_base.err
python -c "
import json
d=json.load(open('gpurun_out/s44_rtm.json')); print('rtm', round(d['value'],1), round(d['ms_per_step'],2))
d=json.load(open('gpurun_out/s44_base.json')); print('baseline', round(d['value'],1), round(d['ms_per_step'],3), 'frac', round(d['roofline']['frac'],3))"
