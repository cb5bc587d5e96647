#!/bin/bash
mkdir -p gpurun_out
UAVDET_MEMBOUND_ONCE=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:dwdynconv -c 2 -f -o gpurun_out/s37_dwdyn python tools/run_membound_kernels.py > gpurun_out/s37_ncu.log 2>&1
ls -la gpurun_out/s37_dwdyn.ncu-rep
