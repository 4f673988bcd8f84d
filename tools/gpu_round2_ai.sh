#!/bin/bash
# bf16 operand mode on the row-streaming kernels + block-tail node in bf16 mode: parity subset, full suite, benches
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x -k "tensor_core_conv_matches_oracle or block_tail" > gpurun_out/r2ai_sub.log 2>&1; echo "subset rc=$?" >> gpurun_out/r2ai_sub.log
tail -4 gpurun_out/r2ai_sub.log
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r2ai_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2ai_tests.log
tail -4 gpurun_out/r2ai_tests.log
timeout 300 python bench.py --workload e20_bf16 --steps 10 --warmup 3 > gpurun_out/r2ai_bench_e20_bf16.json 2> gpurun_out/r2ai_bench_e20_bf16.err
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2ai_bench_e24.json 2> gpurun_out/r2ai_bench_e24.err
for w in e20_bf16 e24; do python -c "
import json
try:
    d=json.load(open('gpurun_out/r2ai_bench_$w.json')); print('$w', d.get('ms_per_step'), d.get('value'), d.get('e2e',{}).get('value'), d['roofline']['kernel'], d['roofline']['frac'])
except Exception as e: print('$w', 'FAILED', e)"; done
