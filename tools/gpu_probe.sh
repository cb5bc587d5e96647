#!/bin/bash
# Runs the GPU kernel tests group by group (each in its own process so a faulting kernel cannot
# poison the CUDA context of the others) and leaves logs in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
i=0
for grp in "nms" "decode" "conv_fwd_plain" "pack_weight or affine or stats or strided or head or per_sample or space_to_depth" "dgrad" "wgrad" "stem or bn_act or upsample or attention or sgd"; do
  i=$((i+1))
  timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "$grp" --tb=short -p no:cacheprovider > gpurun_out/probe_$i.log 2>&1
  echo "group $i [$grp] exit $?" | tee -a gpurun_out/probe_summary.txt
  tail -n 3 gpurun_out/probe_$i.log | tee -a gpurun_out/probe_summary.txt
done
