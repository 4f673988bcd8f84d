#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r2s_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2s_tests.log
timeout 400 python bench.py --workload infonce_sweep > gpurun_out/r2s_sweep.json 2> gpurun_out/r2s_sweep.err
timeout 300 python bench.py --workload e20_bf16 --steps 10 --warmup 3 > gpurun_out/r2s_e20.json 2> gpurun_out/r2s_e20.err
tail -3 gpurun_out/r2s_tests.log; grep -a "FAILED\|^E  " gpurun_out/r2s_tests.log | head
python -c "
import json
d=json.load(open('gpurun_out/r2s_sweep.json')); print(d['value'], d['config']['best_point'])
for p in d['sweep']:
    if p['reg'] or p['candidates']>=2048: print(p)
d=json.load(open('gpurun_out/r2s_e20.json')); print('e20', d['ms_per_step'], d['value'])"
