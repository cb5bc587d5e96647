#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_kernels.py -x -q -k "bn_act" 2>&1 | tail -3
for m in 1 3; do
  UAVDET_BN_REDUCE_MINB=$m UAVDET_BENCH_TIMELINE=1 UAVDET_BENCH_DEBUG=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/s15_minb$m.json 2> gpurun_out/s15_minb$m.txt
done
for m in 1 3 1 3; do
  UAVDET_BN_REDUCE_MINB=$m python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('minb $m', round(d['value'],1), round(d['ms_per_step'],3))"
done
