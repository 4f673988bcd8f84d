#!/bin/bash
# saved outer-ReLU bit mask in the fused batch norm: parity subset + A/B of the e24 step
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -q -x -k "bn_ or block_tail or e24 or scalogram or graph or pool or conv" > gpurun_out/r2ab_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2ab_tests.log
timeout 200 python bench.py --steps 20 --warmup 5 > gpurun_out/r2ab_bench_mask.json 2> gpurun_out/r2ab_bench_mask.err
CPC_NO_EARLY_CROP=1 timeout 200 python bench.py --steps 20 --warmup 5 > gpurun_out/r2ab_bench_nomask.json 2> gpurun_out/r2ab_bench_nomask.err
tail -5 gpurun_out/r2ab_tests.log
for w in mask nomask; do python -c "
import json
try:
    d=json.load(open('gpurun_out/r2ab_bench_$w.json')); print('$w', d.get('ms_per_step'), d.get('value'), d.get('e2e',{}).get('value'))
except Exception as e: print('$w', 'FAILED', e)"; done
