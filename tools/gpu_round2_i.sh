#!/bin/bash
# 1-GPU session: full GPU suite, every bench workload, ncu capture of the dominant kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r2i_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2i_tests.log
for w in e24 raw_wave e20_bf16 long_context infonce_sweep; do
  timeout 400 python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/r2i_bench_$w.json 2> gpurun_out/r2i_bench_$w.err
done
timeout 200 python bench.py --workload raw_wave --batch 64 --steps 10 --warmup 3 > gpurun_out/r2i_bench_raw_wave_b64.json 2> gpurun_out/r2i_bench_raw_wave_b64.err
KEY="cpc_conv_dgrad b64 128x63x156->128x34x156 k30x1 s1x1 [tall_conv_tcgen05_128ch]"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tall128_conv_kernel -c 2 -o gpurun_out/r2i_dominant python tools/profile_kernel.py "$KEY" 1 > gpurun_out/r2i_ncu.log 2>&1
tail -3 gpurun_out/r2i_tests.log
for w in e24 raw_wave e20_bf16 long_context infonce_sweep raw_wave_b64; do python -c "
import json,sys
try:
    d=json.load(open('gpurun_out/r2i_bench_$w.json')); print('$w', d.get('ms_per_step'), d.get('value'), d.get('e2e',{}).get('value'))
except Exception as e: print('$w', 'FAILED', e)
"; done
