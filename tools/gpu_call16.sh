#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py -x -q -k "bn or loss or train or graphed" 2>&1 | tail -4 > gpurun_out/r02_pytest_c16.log
for cfg in "0 0" "1 1" "0 1" "1 0"; do
  set -- $cfg
  UAVDET_BN_FWD_REVERSE=$1 UAVDET_BN_APPLY_REVERSE=$2 UAVDET_BENCH_DEBUG=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline \
     > gpurun_out/r02_c16_bench_f$1_a$2.json 2> gpurun_out/r02_c16_table_f$1_a$2.txt
done
python bench.py --model dysoem --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_c16_bench_dysoem.json 2> /dev/null
for f in gpurun_out/r02_c16_table_*.txt; do echo $f; python - "$f" <<'PY'
import re, sys, collections
agg = collections.defaultdict(float)
for l in open(sys.argv[1]):
    m = re.match(r'\[convtimer\] (\d+) (\w+) (\d+) us', l)
    if m: agg[m.group(2)] += int(m.group(3))
print({k: round(v / 1000, 2) for k, v in agg.items()})
PY
done
for f in gpurun_out/r02_c16_bench_*.json; do python -c "
import json; d=json.load(open('$f')); print('$f'.split('bench_')[1], round(d['value'],1), round(d['ms_per_step'],2))"; done
cat gpurun_out/r02_pytest_c16.log
