#!/bin/bash
# timeline of the graphed step with the weight gradients on their side stream + the two-stream probe
mkdir -p gpurun_out
UAVDET_BENCH_TIMELINE=1 UAVDET_BENCH_DEBUG=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/s9_timeline.json 2> gpurun_out/s9_timeline.txt
python tools/overlap_probe.py > gpurun_out/s9_overlap_probe.txt 2>&1
tail -3 gpurun_out/s9_overlap_probe.txt
python -c "
import json; d=json.load(open('gpurun_out/s9_timeline.json')); print(round(d['value'],1), round(d['ms_per_step'],2))"
