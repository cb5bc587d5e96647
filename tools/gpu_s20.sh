#!/bin/bash
mkdir -p gpurun_out
for v in 3 5 6; do
UAVDET_BN_APPLY_VARIANT=$v python -m pytest tests/test_gpu_kernels.py -x -q -k "bn_act" 2>&1 | tail -1
done
run() { python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['value'],1), round(d['ms_per_step'],3))"; }
UAVDET_BN_APPLY_VARIANT=3 run "apply v3 (regs U2)"
UAVDET_BN_APPLY_VARIANT=5 run "apply v5 (smem U4)"
UAVDET_BN_APPLY_VARIANT=6 run "apply v6 (smem U8)"
UAVDET_BN_APPLY_VARIANT=5 UAVDET_BN_REDUCE_BPS=3 run "apply v5 + reduce bps3"
UAVDET_BN_APPLY_VARIANT=3 UAVDET_BN_REDUCE_BPS=3 run "apply v3 + reduce bps3"
UAVDET_BN_APPLY_VARIANT=3 run "apply v3 again"
UAVDET_BN_APPLY_VARIANT=5 UAVDET_BENCH_TIMELINE=1 UAVDET_BENCH_DEBUG=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/s20_v5.json 2> gpurun_out/s20_v5.txt
UAVDET_BN_APPLY_VARIANT=3 UAVDET_BENCH_TIMELINE=1 UAVDET_BENCH_DEBUG=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/s20_v3.json 2> gpurun_out/s20_v3.txt
