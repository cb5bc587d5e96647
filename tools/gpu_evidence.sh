#!/bin/bash
# final ncu evidence of round 2 (one gpurun call; every command runs plain first).  The .ncu-rep files are summarised ON
# THE BOX and deleted: gpurun brings back at most 64 MiB.
mkdir -p gpurun_out
python bench.py --eager --profile-step > gpurun_out/r02_ev_plain_step.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off \
   --csv --log-file gpurun_out/r02_launches_train_step_b32.csv python bench.py --eager --profile-step > gpurun_out/r02_ev_ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:igemm_kernel -s 18 -c 8 -f \
   -o /tmp/r02_igemm_in_step python bench.py --eager --profile-step > gpurun_out/r02_ev_ncu2.log 2>&1
python tools/ncu_summary.py /tmp/r02_igemm_in_step.ncu-rep > gpurun_out/r02_ncu_full_igemm_in_train_step.txt 2>&1
ncu -i /tmp/r02_igemm_in_step.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv, sys
rows = list(csv.reader(sys.stdin)); hdr = rows[0]
keep = [i for i, h in enumerate(hdr) if h in ('Kernel Name', 'Grid Size') or 'pipe_tensor' in h or h.startswith('gpu__time_duration') or 'sm__throughput' in h]
for r in rows[:2] + rows[2:]:
    print(' | '.join(r[i][:60] for i in keep))
" > gpurun_out/r02_ncu_tensor_pipe_metrics_igemm.txt 2>&1
for i in 1 4; do python tools/ncu_top_stalls.py /tmp/r02_igemm_in_step.ncu-rep $i 12 >> gpurun_out/r02_ncu_full_igemm_in_train_step.txt 2>&1; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
   --log-file gpurun_out/r02_launches_dyyolo_step_b32.csv python tools/profile_step_launches.py dyyolo > gpurun_out/r02_ev_ncu3.log 2>&1
UAVDET_MEMBOUND_ONCE=1 timeout 600 ncu --set full --clock-control none \
  -k regex:'dwdynconv|gn_stats|gn_apply|bilinear2x|decode_yolo|rtm_head_post|gap_kernel|gap_nchw|encode_targets|sgd_momentum|bn_act_fwd|bn_bwd|upsample2x_fwd|cxcywh|stem1x1|stem_s2d|head_grad_pack' \
  -o /tmp/r02_membound_final -f python tools/run_membound_kernels.py > gpurun_out/r02_ev_ncu4.log 2>&1
python tools/ncu_membound_table.py /tmp/r02_membound_final.ncu-rep > gpurun_out/r02_ncu_full_membound_kernels_final.txt 2>&1
python tools/run_membound_kernels.py > gpurun_out/r02_membound_kernels_event_timed_final.jsonl 2> gpurun_out/r02_ev_membound.err
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_ev_bench_full.json 2> gpurun_out/r02_ev_bench_full.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_ev_bench_reference.json 2> gpurun_out/r02_ev_bench_reference.err
python bench.py --impl torch-gpu --steps 10 --warmup 3 > gpurun_out/r02_ev_bench_torch_gpu.json 2> gpurun_out/r02_ev_bench_torch_gpu.err
du -sh gpurun_out
for m in dyyolo dysoem rtm-infer; do
  python bench.py --model $m --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_ev_bench_$m.json 2> gpurun_out/r02_ev_bench_$m.err
done
python tools/bench_infer.py > gpurun_out/r02_ev_inference_configs.jsonl 2> gpurun_out/r02_ev_inference.err
ls -la gpurun_out | tail -30
