#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r2z_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2z_tests.log
timeout 400 python bench.py --workload e29 --steps 5 --warmup 3 > gpurun_out/r2z_bench_e29.json 2> gpurun_out/r2z_bench_e29.err
tail -3 gpurun_out/r2z_tests.log; grep -a "FAILED\|^E  " gpurun_out/r2z_tests.log | head
python -c "
import json
d=json.load(open('gpurun_out/r2z_bench_e29.json')); print('e29', d['ms_per_step'], d['value'], d['e2e']['value'])
for f in d['kernel_families'][:8]: print({k:(round(v,3) if isinstance(v,float) else v) for k,v in f.items()})" || tail -5 gpurun_out/r2z_bench_e29.err
