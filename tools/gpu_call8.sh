#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r02_pytest_c8.log
for m in baseline dyyolo dysoem rtm-infer; do
  python bench.py --model $m --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_c8_bench_$m.json 2> gpurun_out/r02_c8_bench_$m.err
done
python tools/run_membound_kernels.py > gpurun_out/r02_c8_membound.jsonl 2>&1
tail -4 gpurun_out/r02_pytest_c8.log
