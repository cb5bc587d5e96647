#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_kernels.py -x -q -k "dwdynconv or rtm or bilinear" 2>&1 | tail -5 > gpurun_out/r02_pytest_dw.log
python -m pytest tests/test_gpu_models.py tests/test_gpu_real_shapes.py -x -q -k "rtm" 2>&1 | tail -5 >> gpurun_out/r02_pytest_dw.log
python tools/run_membound_kernels.py 2>&1 | grep dwdynconv > gpurun_out/r02_dw_membound.jsonl
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 tools/dp_overlap_trace.py > gpurun_out/r02_dp_overlap_timeline.txt 2> gpurun_out/r02_dp_overlap.err
python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_dpb_n1.json 2> gpurun_out/r02_dpb_n1.err
for cfg in "8 8" "8 -1" "4 8" "8 16" "0 0"; do
  set -- $cfg
  UAVDET_DP_SM_MARGIN=$1 UAVDET_DP_SM_MARGIN_HOLD=$2 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 \
    bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_dpb_n2_m$1_h$2.json 2> gpurun_out/r02_dpb_n2_m$1_h$2.err
done
for f in gpurun_out/r02_dpb_*.json; do python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f'.split('dpb_')[1], d['n_gpus'], round(d['value'],1), round(d['ms_per_step'],2))"; done
cat gpurun_out/r02_pytest_dw.log
