#!/bin/bash
# GELU / SiLU instances of the AFFINE epilogue: tests + RTMUAVDet inference bench
mkdir -p gpurun_out
python -m pytest tests/ -m gpu -x -q 2>&1 | tail -3 > gpurun_out/s25_tests.log; cat gpurun_out/s25_tests.log
python bench.py --model rtm-infer > gpurun_out/s25_rtm.json 2> gpurun_out/s25_rtm.err; python -c "
import json; d=json.load(open('gpurun_out/s25_rtm.json')); print('rtm', round(d['value'],1), round(d['ms_per_step'],2))"
