#!/bin/bash
mkdir -p gpurun_out
python tools/bench_nms_floor.py > gpurun_out/s19_nms_floor.txt 2>&1
cd multimodal_uav_det_b200 && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -DUAVDET_NMS_PROFILE -c csrc/nms.cu -o build/nms.o && nvcc -shared -o libuavdet_b200.so build/*.o -gencode arch=compute_100a,code=sm_100a && cd ..
python tools/bench_nms_floor.py 2>&1 | grep -v "^floor" | sort | uniq -c | sort -rn | head -20 > gpurun_out/s19_nms_floor_profile.txt
cat gpurun_out/s19_nms_floor.txt; head -12 gpurun_out/s19_nms_floor_profile.txt
