#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py -x -q -k "bn_act or conv_fwd or conv_dgrad or graphed or train_step" 2>&1 | tail -4
for m in 1 0 1 0; do
  UAVDET_NO_PDL=$m python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2> gpurun_out/s16_err_$m.txt | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('no_pdl $m', round(d['value'],1), round(d['ms_per_step'],3), d['config'].get('final_loss'))"
done
tail -3 gpurun_out/s16_err_0.txt
