#!/bin/bash
# final state of the round: the three driver entry points, plain
mkdir -p gpurun_out
python -m pytest tests/ -m gpu -x -q 2>&1 | tail -3 > gpurun_out/final_tests.log; cat gpurun_out/final_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py > gpurun_out/final_bench_default.json 2> gpurun_out/final_bench_default.err; python -c "
import json; d=json.load(open('gpurun_out/final_bench_default.json')); print(round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],3), 'traffic', d['roofline']['traffic'], 'launches', d['gpu_launches'], d['clocks'], d['cpu_baseline']['value'])"
python bench.py --model rtm-infer --no-cpu-baseline > gpurun_out/final_bench_rtm.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/final_bench_rtm.json')); print('rtm', round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1))"
