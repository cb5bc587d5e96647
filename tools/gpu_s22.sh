#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/ -m gpu -x -q -k "dyn or dysoem or dyyolo or host or symbol" 2>&1 | tail -4
python bench.py --model dysoem --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('dysoem', round(d['value'],1), round(d['ms_per_step'],3), d['config'].get('final_loss'))"
