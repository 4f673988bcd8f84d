#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do
timeout 300 python bench.py --steps 30 --warmup 5 > gpurun_out/r2r_bench_new$i.json 2> gpurun_out/r2r_bench_new$i.err
CPC_NO_MMA_SMALL_WGRAD=1 timeout 300 python bench.py --steps 30 --warmup 5 > gpurun_out/r2r_bench_old$i.json 2> gpurun_out/r2r_bench_old$i.err
done
python -c "
import json
for n in ('new1','old1','new2','old2'):
    d=json.load(open('gpurun_out/r2r_bench_%s.json'%n)); print(n, round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), d['clocks']['sm_mhz'], d['clocks']['reasons'])"
