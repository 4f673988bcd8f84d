#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r2v_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2v_tests.log
timeout 400 python bench.py --workload infonce_sweep > gpurun_out/r2v_sweep.json 2> gpurun_out/r2v_sweep.err
tail -3 gpurun_out/r2v_tests.log; grep -a "FAILED\|^E  " gpurun_out/r2v_tests.log | head
python -c "
import json
d=json.load(open('gpurun_out/r2v_sweep.json')); print(d['value'], d['config']['best_point'])
for p in d['sweep']:
    if p['candidates']>=8192: print(p['mode'], p['K'], p['score'], p['reg'], round(p['ms'],2), round(p['tflops'],1))"
