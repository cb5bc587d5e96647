#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_kernels.py -x -q -k "umma_descriptor" 2>&1 | tail -30 > gpurun_out/r02_probe.log
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py -x -q 2>&1 | tail -15 > gpurun_out/r02_pytest_c10.log
python tools/run_membound_kernels.py 2>&1 | grep dwdynconv > gpurun_out/r02_dw2_membound.jsonl
UAVDET_WGRAD_2CTA=0 python tools/bench_conv_layers.py --ours-only > gpurun_out/r02_layers_wg1.json 2> /dev/null
UAVDET_WGRAD_2CTA=1 python tools/bench_conv_layers.py --ours-only > gpurun_out/r02_layers_wg2.json 2> /dev/null
UAVDET_WGRAD_2CTA=0 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_c10_bench_wg1.json 2> gpurun_out/r02_c10_bench_wg1.err
UAVDET_WGRAD_2CTA=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_c10_bench_wg2.json 2> gpurun_out/r02_c10_bench_wg2.err
tail -3 gpurun_out/r02_pytest_c10.log; tail -3 gpurun_out/r02_probe.log
