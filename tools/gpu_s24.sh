#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_real_shapes.py -x -q -k "conv_fwd or rtm or affine or conv_dgrad" 2>&1 | tail -2
UAVDET_BENCH_DEBUG=1 python bench.py --model rtm-infer --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/s24_rtm.json 2> gpurun_out/s24_rtm_table.txt
UAVDET_IGEMM_EPC_BIAS=1 UAVDET_BENCH_DEBUG=1 python bench.py --model rtm-infer --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/s24_rtm_bias.json 2> gpurun_out/s24_rtm_bias_table.txt
for f in gpurun_out/s24_rtm.json gpurun_out/s24_rtm_bias.json; do python -c "
import json; d=json.load(open('$f')); print('$f', round(d['value'],1), round(d['ms_per_step'],3))"; done
