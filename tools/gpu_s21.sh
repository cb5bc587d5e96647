#!/bin/bash
# final-code sanity of the data-parallel path at N = 2 (all three training configurations + replicas of the inference one)
mkdir -p gpurun_out
python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/s21_n1_baseline.json 2>/dev/null
for m in baseline dyyolo dysoem rtm-infer; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 10 --warmup 3 --model $m --no-cpu-baseline > gpurun_out/s21_n2_$m.json 2> gpurun_out/s21_n2_$m.err
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/s21_n2_reference.json 2> gpurun_out/s21_n2_reference.err
for f in gpurun_out/s21_*.json; do python -c "
import json
for l in open('$f'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('$f', d.get('n_gpus'), round(d.get('value',0),1), round(d.get('ms_per_step',0),2), d.get('config',{}).get('final_loss'))"; done
tail -3 gpurun_out/s21_n2_baseline.err
