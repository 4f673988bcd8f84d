#!/bin/bash
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -q -x -k "bn_ or block_tail or e24 or scalogram or graph" > gpurun_out/r2ah_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2ah_tests.log
tail -4 gpurun_out/r2ah_tests.log
timeout 200 python bench.py --steps 20 --warmup 5 > gpurun_out/r2ah_bench.json 2> gpurun_out/r2ah_bench.err
python -c "
import json
d=json.load(open('gpurun_out/r2ah_bench.json')); print(d.get('ms_per_step'), d.get('value'), d['e2e']['value'])
for k in d['kernels']:
    if 'bn_relu' in k['key'] or 'pool' in k['key']: print('   ', k['key'], round(k['avg_ms'],4))"
