#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r2j_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2j_tests.log
for w in raw_wave e20_bf16; do
  timeout 300 python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/r2j_bench_$w.json 2> gpurun_out/r2j_bench_$w.err
done
timeout 200 python bench.py --workload raw_wave --batch 64 --steps 10 --warmup 3 > gpurun_out/r2j_bench_raw_wave_b64.json 2> gpurun_out/r2j_bench_raw_wave_b64.err
tail -3 gpurun_out/r2j_tests.log; grep -a "FAILED\|^E  " gpurun_out/r2j_tests.log | head -20
for w in raw_wave e20_bf16 raw_wave_b64; do python -c "
import json,sys
try:
    d=json.load(open('gpurun_out/r2j_bench_$w.json')); print('$w', d.get('ms_per_step'), d.get('value'), d.get('e2e',{}).get('value'), d['roofline']['kernel'], d['roofline']['frac'])
except Exception as e: print('$w', 'FAILED', e)
"; done
