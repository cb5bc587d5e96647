#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py -x -q 2>&1 | tail -15 > gpurun_out/r02_pytest_c7.log
for v in 0 1 2 3; do
  UAVDET_BN_APPLY_VARIANT=$v python tools/run_membound_kernels.py 2>&1 | grep -E "bn_bwd|bilinear|gap_kernel" > gpurun_out/r02_membound_bnvar$v.jsonl
done
for v in 0 3; do
  UAVDET_BN_APPLY_VARIANT=$v python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_bnvar$v.json 2> gpurun_out/r02_bench_bnvar$v.err
done
tail -4 gpurun_out/r02_pytest_c7.log
