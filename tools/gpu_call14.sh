#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r02_pytest_c14.log
for m in dysoem rtm-infer dyyolo; do
  python bench.py --model $m --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_c14_bench_$m.json 2> gpurun_out/r02_c14_bench_$m.err
done
python tools/bench_infer.py > gpurun_out/r02_c14_infer.jsonl 2> gpurun_out/r02_c14_infer.err
tail -3 gpurun_out/r02_pytest_c14.log
