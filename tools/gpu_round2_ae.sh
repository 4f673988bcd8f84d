#!/bin/bash
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -q -x -k "pool or e24 or experiment or raw_wave or golden" > gpurun_out/r2ae_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2ae_tests.log
timeout 200 python tools/probes/ar_graph_probe.py > gpurun_out/r2ae_probe.log 2>&1
timeout 200 python bench.py --steps 20 --warmup 5 > gpurun_out/r2ae_bench.json 2> gpurun_out/r2ae_bench.err
tail -4 gpurun_out/r2ae_tests.log
tail -3 gpurun_out/r2ae_probe.log
python -c "
import json
d=json.load(open('gpurun_out/r2ae_bench.json')); print(d.get('ms_per_step'), d.get('value'), d.get('e2e',{}).get('value'))"
