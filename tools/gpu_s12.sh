#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_kernels.py -x -q -k "nms or bn_act" 2>&1 | tail -4 > gpurun_out/s12_tests.log
python -m pytest tests/test_gpu_real_shapes.py tests/test_gpu_models.py -x -q -k "rtm or detect or return_ap" 2>&1 | tail -4 >> gpurun_out/s12_tests.log
cat gpurun_out/s12_tests.log
python tools/bench_stem1x1.py > gpurun_out/s12_stem1x1.txt 2>&1; cat gpurun_out/s12_stem1x1.txt
UAVDET_BENCH_DEBUG=1 python bench.py --model rtm-infer --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/s12_rtm.json 2> gpurun_out/s12_rtm_table.txt
python bench.py --model dysoem --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/s12_dysoem.json 2>/dev/null
for f in gpurun_out/s12_*.json; do python -c "
import json; d=json.load(open('$f')); print('$f', round(d['value'],1), round(d['ms_per_step'],3))"; done
