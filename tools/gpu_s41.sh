#!/bin/bash
# --set full of the first igemm launches of the training step (3x3 32->64 stride 2, 1x1 64->32, the pixel-pair 3x3 32->64) and of the
# pair data gradients at the end of backward
mkdir -p gpurun_out
python bench.py --eager --profile-step > gpurun_out/s41_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:igemm_kernel -c 3 -f \
   -o /tmp/s41_pair python bench.py --eager --profile-step > gpurun_out/s41_ncu.log 2>&1
python tools/ncu_summary.py /tmp/s41_pair.ncu-rep > gpurun_out/r02_ncu_full_igemm_pair_32_64_3x3_320.txt 2>&1
tail -40 gpurun_out/r02_ncu_full_igemm_pair_32_64_3x3_320.txt
