#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/r2o_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2o_tests.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2o_bench_overlap.json 2> gpurun_out/r2o_bench_overlap.err
CPC_NO_BWD_OVERLAP=1 timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2o_bench_serial.json 2> gpurun_out/r2o_bench_serial.err
tail -3 gpurun_out/r2o_tests.log
python -c "
import json
for n in ('overlap','serial'):
    d=json.load(open('gpurun_out/r2o_bench_%s.json'%n)); print(n, d['ms_per_step'], d['value'], d['e2e']['ms_per_step'])"
