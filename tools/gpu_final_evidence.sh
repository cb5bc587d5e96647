#!/bin/bash
# End-of-round evidence from the final binaries (one gpurun call; every program runs plain before it runs under ncu).
mkdir -p gpurun_out
O=gpurun_out
python bench.py --steps 20 --warmup 5 > $O/r02_final_bench_full.json 2> $O/ev_bench_full.err
python bench.py --model rtm-infer --steps 10 --warmup 3 --no-cpu-baseline > $O/r02_final_bench_rtm-infer.json 2> $O/ev_rtm.err
python bench.py --model dyyolo --steps 10 --warmup 3 --no-cpu-baseline > $O/r02_final_bench_dyyolo.json 2> $O/ev_dyyolo.err
python bench.py --model dysoem --steps 10 --warmup 3 --no-cpu-baseline > $O/r02_final_bench_dysoem.json 2> $O/ev_dysoem.err
python bench.py --impl torch-gpu --steps 10 --warmup 3 > $O/r02_final_bench_torch_gpu.json 2> $O/ev_torch.err
UAVDET_BENCH_DEBUG=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/ev_dbg.json 2> $O/r02_final_conv_launch_table_in_graph.txt
UAVDET_BENCH_DEBUG=1 python bench.py --model rtm-infer --steps 10 --warmup 3 --no-cpu-baseline > $O/ev_dbg_rtm.json 2> $O/r02_final_conv_launch_table_rtm-infer.txt
# launch lists
python bench.py --eager --profile-step > $O/ev_plain_step.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off \
   --csv --log-file $O/r02_launches_train_step_b32_with_dram_bytes.csv python bench.py --eager --profile-step > $O/ev_ncu1.log 2>&1
python tools/conv_traffic_from_ncu.py $O/r02_launches_train_step_b32_with_dram_bytes.csv > $O/r02_conv_dram_traffic_per_step.json
python tools/profile_step_launches.py rtm-infer > $O/ev_plain_rtm.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv \
   --log-file $O/r02_launches_rtm-infer_step.csv python tools/profile_step_launches.py rtm-infer > $O/ev_ncu2.log 2>&1
python - <<'PY' > gpurun_out/r02_launch_summary_rtm-infer_step.txt
import collections, csv, re
lines = [l for l in open('gpurun_out/r02_launches_rtm-infer_step.csv') if not l.startswith('==')]
agg = collections.defaultdict(lambda: [set(), 0.0, 0.0])
for r in csv.DictReader(lines):
    v = float(r['Metric Value'].replace(',', '')); u = r['Metric Unit']
    name = re.sub(r'^void ', '', re.sub(r'\(.*', '', r['Kernel Name']))[:64]
    a = agg[name]; a[0].add(r['ID'])
    if r['Metric Name'].startswith('gpu__time'):
        a[1] += v / 1e6 if u in ('ns', 'nsecond') else v / 1e3 if u in ('us', 'usecond') else v
    else:
        a[2] += v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[u]
tot = sum(a[1] for a in agg.values())
print(f'total kernel time {tot:.3f} ms over {sum(len(a[0]) for a in agg.values())} launches (cold, serialised: shares, not absolutes)')
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:28]:
    print(f'{a[1]:9.3f} ms {100 * a[1] / tot:5.1f}%  x{len(a[0]):4d}  {a[2] / 1e6:9.1f} MB DRAM  {a[2] / 1e3 / max(a[1], 1e-9) / 1e3:7.0f} GB/s  {k}')
PY
python tools/launch_summary.py $O/r02_launches_train_step_b32_with_dram_bytes.csv 30 > $O/ev_summary_baseline_raw.txt 2>&1
# --set full of the kernels this session changed, inside the RTMUAVDet step
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'dwdynconv|gn_stats' -c 12 -f \
   -o /tmp/r02_rtm_stream python tools/profile_step_launches.py rtm-infer > $O/ev_ncu3.log 2>&1
python tools/ncu_membound_table.py /tmp/r02_rtm_stream.ncu-rep > $O/r02_ncu_full_dwdynconv_in_rtm_step.txt 2>&1
ls -la $O | tail -25
