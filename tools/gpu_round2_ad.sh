#!/bin/bash
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -q -x -k "e24 or experiment or raw_wave or golden or nccl" > gpurun_out/r2ad_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2ad_tests.log
timeout 200 python tools/probes/ar_graph_probe.py > gpurun_out/r2ad_probe_on.log 2>&1
CPC_NO_AR_FUSION=1 timeout 200 python tools/probes/ar_graph_probe.py > gpurun_out/r2ad_probe_off.log 2>&1
timeout 200 python bench.py --steps 20 --warmup 5 > gpurun_out/r2ad_bench_on.json 2> gpurun_out/r2ad_bench_on.err
CPC_NO_AR_FUSION=1 timeout 200 python bench.py --steps 20 --warmup 5 > gpurun_out/r2ad_bench_off.json 2> gpurun_out/r2ad_bench_off.err
tail -4 gpurun_out/r2ad_tests.log
echo on; tail -3 gpurun_out/r2ad_probe_on.log; echo off; tail -3 gpurun_out/r2ad_probe_off.log
for w in on off; do python -c "
import json
try:
    d=json.load(open('gpurun_out/r2ad_bench_$w.json')); print('$w', d.get('ms_per_step'), d.get('value'), d.get('e2e',{}).get('value'))
except Exception as e: print('$w', 'FAILED', e)"; done
