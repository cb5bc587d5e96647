#!/bin/bash
for m in 0 1 14 2 12 3 15 0; do
  UAVDET_PDL_MASK=$m python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('pdl_mask $m', round(d['value'],1), round(d['ms_per_step'],3))"
done
