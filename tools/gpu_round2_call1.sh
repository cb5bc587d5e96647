#!/bin/bash
# first GPU call of round 2: library bars + evidence for the memory-bound kernels (outputs under gpurun_out/)
set -x
mkdir -p gpurun_out
python -c "import psutil,os; print('ram_gib', psutil.virtual_memory().total/2**30, 'cpus', os.cpu_count())" > gpurun_out/r02_probe.txt 2>&1
nvidia-smi -L >> gpurun_out/r02_probe.txt
python bench.py --impl torch-gpu --steps 10 --warmup 3 > gpurun_out/r02_torch_gpu.json 2> gpurun_out/r02_torch_gpu.err
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_ours_before.json 2> gpurun_out/r02_ours_before.err
python tools/bench_conv_layers.py > gpurun_out/r02_conv_layers.json 2> gpurun_out/r02_conv_layers.err
python tools/bench_nms_vs_torchvision.py > gpurun_out/r02_nms_vs_tv.jsonl 2> gpurun_out/r02_nms_vs_tv.err
python tools/run_membound_kernels.py > gpurun_out/r02_membound.jsonl 2> gpurun_out/r02_membound.err
UAVDET_MEMBOUND_ONCE=1 timeout 600 ncu --set full --clock-control none --import-source on \
  -k regex:'dwdynconv|gn_stats|gn_apply|bilinear2x|decode_yolo|rtm_head_post|gap_kernel|encode_targets|sgd_momentum|bn_act_fwd|bn_bwd|upsample2x_fwd|cxcywh' \
  -o gpurun_out/r02_membound -f python tools/run_membound_kernels.py > gpurun_out/r02_membound_ncu.log 2>&1
echo done
