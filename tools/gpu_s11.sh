#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_kernels.py -x -q -k "bn_act or stem or adaptive" 2>&1 | tail -4 > gpurun_out/s11_tests.log
python -m pytest tests/test_gpu_real_shapes.py tests/test_gpu_models.py -x -q -k "adaptive or dysoem" 2>&1 | tail -4 >> gpurun_out/s11_tests.log
cat gpurun_out/s11_tests.log
for f in 0 1; do
  if [ $f = 1 ]; then export UAVDET_BN_FUSE_FWD=1; fi
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/s11_bench_fuse$f.json 2>/dev/null
  python bench.py --model dysoem --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/s11_dysoem_fuse$f.json 2>/dev/null
done
for f in gpurun_out/s11_*.json; do python -c "
import json; d=json.load(open('$f')); print('$f', round(d['value'],1), round(d['ms_per_step'],3))"; done
