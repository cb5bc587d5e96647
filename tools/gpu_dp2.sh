#!/bin/bash
# 2-GPU data-parallel A/B: SM margin for the collective (0 = old behaviour but with per-layer hooks, 8, 16)
mkdir -p gpurun_out
python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_dp_n1.json 2> gpurun_out/r02_dp_n1.err
for m in 8 0 16; do
  UAVDET_DP_SM_MARGIN=$m NCCL_DEBUG=${NCCL_DBG:-WARN} python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_dp_n2_margin$m.json 2> gpurun_out/r02_dp_n2_margin$m.err
done
grep -h '"metric"' gpurun_out/r02_dp_n*.json | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['n_gpus'], round(d['value'],1), round(d['ms_per_step'],2))
"
