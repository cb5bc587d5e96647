#!/bin/bash
# One 8-GPU box: the 1 -> N tables of BASELINE.json's configurations.  Smaller-N jobs share the box on disjoint GPUs.
mkdir -p gpurun_out
run() {  # run <gpus csv> <port> <n> <model> <tag> [extra env]
  local gpus=$1 port=$2 n=$3 model=$4 tag=$5
  if [ "$n" = "1" ]; then
    CUDA_VISIBLE_DEVICES=$gpus python bench.py --gpus 1 --model $model --steps 10 --warmup 3 --no-cpu-baseline \
      > gpurun_out/r02_scale_${tag}.json 2> gpurun_out/r02_scale_${tag}.err
  else
    CUDA_VISIBLE_DEVICES=$gpus python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus $n --model $model --steps 10 --warmup 3 > gpurun_out/r02_scale_${tag}.json 2> gpurun_out/r02_scale_${tag}.err
  fi
}
python -c "from multimodal_uav_det_b200 import build; build.build_library()"
# phase 1
run 0,1,2,3 29501 4 baseline baseline_n4 & run 4,5 29502 2 baseline baseline_n2 & run 6,7 29503 2 dysoem dysoem_n2 & wait
# phase 2
run 0,1,2,3 29504 4 dysoem dysoem_n4 & run 4 0 1 rtm-infer rtm_n1 & run 5 0 1 baseline baseline_n1 & run 6 0 1 dyyolo dyyolo_n1 & run 7 0 1 dysoem dysoem_n1 & wait
# phases 3-6: the whole box
NCCL_DEBUG=INFO run 0,1,2,3,4,5,6,7 29505 8 baseline baseline_n8
grep -E "NCCL INFO (Connected|Channel|comm|NVLS|Using|ncclCommInitRank)" gpurun_out/r02_scale_baseline_n8.err | head -60 > gpurun_out/r02_scale_nccl_rank_log.txt
run 0,1,2,3,4,5,6,7 29506 8 dyyolo dyyolo_n8
run 0,1,2,3,4,5,6,7 29507 8 dysoem dysoem_n8
run 0,1,2,3,4,5,6,7 29508 8 rtm-infer rtm_n8
for f in gpurun_out/r02_scale_*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split("r02_scale_")[1], d["n_gpus"], round(d["value"], 1), round(d["ms_per_step"], 2), round(d["e2e"]["value"], 1))
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
