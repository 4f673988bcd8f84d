#!/bin/bash
# per-step InfoNCE on the CUDA-core path: one CTA per (tile, step[, E slice]); parity + raw_wave benches
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x -k "infonce or raw or replay or train or validate" > gpurun_out/r2al_sub.log 2>&1; echo "subset rc=$?" >> gpurun_out/r2al_sub.log
tail -4 gpurun_out/r2al_sub.log
timeout 300 python bench.py --workload raw_wave --steps 10 --warmup 3 > gpurun_out/r2al_bench_raw_wave.json 2> gpurun_out/r2al_bench_raw_wave.err
timeout 300 python bench.py --workload raw_wave --batch 64 --steps 10 --warmup 3 > gpurun_out/r2al_bench_raw_wave_b64.json 2> gpurun_out/r2al_bench_raw_wave_b64.err
for w in raw_wave raw_wave_b64; do python -c "
import json
try:
    d=json.load(open('gpurun_out/r2al_bench_$w.json')); print('$w', d.get('ms_per_step'), d.get('value'), d.get('e2e',{}).get('value'), d['roofline']['kernel'], d['roofline']['frac'])
    print([(k['key'][:60], round(k['avg_ms'],3)) for k in d['kernels'] if 'infonce' in k['key']])
except Exception as e: print('$w', 'FAILED', e)"; done
