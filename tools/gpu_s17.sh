#!/bin/bash
mkdir -p gpurun_out
for ov in 1 0; do for m in 1 0; do
  if [ $ov = 1 ]; then export UAVDET_NO_WGRAD_OVERLAP=1; else unset UAVDET_NO_WGRAD_OVERLAP; fi
  UAVDET_NO_PDL=$m python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('no_overlap $ov no_pdl $m', round(d['value'],1), round(d['ms_per_step'],3))"
done; done
UAVDET_BENCH_TIMELINE=1 UAVDET_BENCH_DEBUG=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/s17_pdl_timeline.json 2> gpurun_out/s17_pdl_timeline.txt
