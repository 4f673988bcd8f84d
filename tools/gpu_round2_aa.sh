#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r2aa_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2aa_tests.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2aa_bench.json 2> gpurun_out/r2aa_bench.err
tail -3 gpurun_out/r2aa_tests.log; grep -a "FAILED\|^E  " gpurun_out/r2aa_tests.log | head
python -c "
import json
d=json.load(open('gpurun_out/r2aa_bench.json')); print(d['ms_per_step'], d['value'], d['metrics']['infonce_fwd_ms'], d['metrics']['infonce_bwd_ms'])"
