#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2x_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2x_smoke.log
tail -3 gpurun_out/r2x_smoke.log
timeout 200 python bench.py --steps 20 --warmup 5 > gpurun_out/r2x_bench.json 2> gpurun_out/r2x_bench.err
python -c "
import json
d=json.load(open('gpurun_out/r2x_bench.json')); print(d['ms_per_step'], d['value'], d['roofline']['traffic'], d['roofline']['traffic_source'], d['gpu_launches'])"
