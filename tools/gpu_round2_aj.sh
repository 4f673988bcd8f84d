#!/bin/bash
# dynamic tile scheduling in the row-streaming conv kernels, AR model under bf16 autocast in bf16 mode
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x -k "tensor_core_conv_matches_oracle or block_tail or tall_conv or bf16_mode" -s > gpurun_out/r2aj_sub.log 2>&1; echo "subset rc=$?" >> gpurun_out/r2aj_sub.log
grep -a "e20 fp32\|cosine" gpurun_out/r2aj_sub.log; tail -3 gpurun_out/r2aj_sub.log
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r2aj_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2aj_tests.log
tail -4 gpurun_out/r2aj_tests.log
timeout 300 python bench.py --workload e20_bf16 --steps 10 --warmup 3 > gpurun_out/r2aj_bench_e20_bf16.json 2> gpurun_out/r2aj_bench_e20_bf16.err
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2aj_bench_e24.json 2> gpurun_out/r2aj_bench_e24.err
timeout 300 python tools/profile_step.py --warmup 3 --steps 2 --table > gpurun_out/r2aj_table.log 2>&1
grep -a "tall_conv_tcgen05" gpurun_out/r2aj_table.log | cut -c1-160
for w in e20_bf16 e24; do python -c "
import json
try:
    d=json.load(open('gpurun_out/r2aj_bench_$w.json')); print('$w', d.get('ms_per_step'), d.get('value'), d.get('e2e',{}).get('value'), d['roofline']['kernel'], d['roofline']['frac'])
except Exception as e: print('$w', 'FAILED', e)"; done
