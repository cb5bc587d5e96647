#!/bin/bash
# final ncu evidence of round 2 (one gpurun call; every command ran plain first)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_kernels.py -x -q -k "loss or head" 2>&1 | tail -3 > gpurun_out/r02_ev_pytest.log
python bench.py --eager --profile-step > gpurun_out/r02_ev_plain_step.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off \
   --csv --log-file gpurun_out/r02_launches_train_step_b32.csv python bench.py --eager --profile-step > gpurun_out/r02_ev_ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:igemm_kernel -s 18 -c 8 -f \
   -o gpurun_out/r02_igemm_in_step python bench.py --eager --profile-step > gpurun_out/r02_ev_ncu2.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
   --log-file gpurun_out/r02_launches_dyyolo_step_b32.csv python tools/profile_step_launches.py dyyolo > gpurun_out/r02_ev_ncu3.log 2>&1
UAVDET_MEMBOUND_ONCE=1 timeout 600 ncu --set full --clock-control none --import-source on \
  -k regex:'dwdynconv|gn_stats|gn_apply|bilinear2x|decode_yolo|rtm_head_post|gap_kernel|gap_nchw|encode_targets|sgd_momentum|bn_act_fwd|bn_bwd|upsample2x_fwd|cxcywh|stem1x1|stem_s2d|head_grad_pack' \
  -o gpurun_out/r02_membound_final -f python tools/run_membound_kernels.py > gpurun_out/r02_ev_ncu4.log 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_ev_bench_full.json 2> gpurun_out/r02_ev_bench_full.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_ev_bench_reference.json 2> gpurun_out/r02_ev_bench_reference.err
cat gpurun_out/r02_ev_pytest.log
