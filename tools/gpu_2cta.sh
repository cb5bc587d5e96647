#!/bin/bash
mkdir -p gpurun_out
UAVDET_IGEMM_2CTA=2 timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "conv or stem or dgrad" 2>&1 | tail -25 > gpurun_out/r02_2cta_pytest_all.log
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py -x -q 2>&1 | tail -25 > gpurun_out/r02_2cta_pytest_default.log
UAVDET_IGEMM_2CTA=0 python tools/bench_conv_layers.py --ours-only > gpurun_out/r02_layers_1cta.json 2> gpurun_out/r02_layers_1cta.err
UAVDET_IGEMM_2CTA=1 python tools/bench_conv_layers.py --ours-only > gpurun_out/r02_layers_2cta.json 2> gpurun_out/r02_layers_2cta.err
UAVDET_IGEMM_2CTA=2 python tools/bench_conv_layers.py --ours-only > gpurun_out/r02_layers_2cta_all.json 2> gpurun_out/r02_layers_2cta_all.err
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_2cta.json 2> gpurun_out/r02_bench_2cta.err
tail -3 gpurun_out/r02_2cta_pytest_all.log gpurun_out/r02_2cta_pytest_default.log
