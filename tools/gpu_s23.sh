#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/ -m gpu -x -q 2>&1 | tail -5
for e in 1 0; do
  if [ $e = 1 ]; then export UAVDET_IGEMM_NO_EPC=1; else unset UAVDET_IGEMM_NO_EPC; fi
  python bench.py --model rtm-infer --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('rtm no_epc=$e', round(d['value'],1), round(d['ms_per_step'],3))"
done
UAVDET_BENCH_DEBUG=1 python bench.py --model rtm-infer --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/s23_rtm.json 2> gpurun_out/s23_rtm_table.txt
python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('baseline', round(d['value'],1), round(d['ms_per_step'],3), d['roofline']['traffic'], d['roofline']['frac'])"
python tools/bench_infer.py 2>/dev/null | cut -c1-600
