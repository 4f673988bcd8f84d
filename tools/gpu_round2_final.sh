#!/bin/bash
# final evidence of round 2: full GPU suite, smoke, every workload's bench line, the reference arm, launch list,
# ncu --set full of the dominant kernel / the CQT / the batch-norm kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2f_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2f_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2f_smoke.log 2>&1
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2f_bench_e24.json 2> gpurun_out/r2f_bench_e24.err
for w in raw_wave e20_bf16 long_context infonce_sweep; do
  timeout 400 python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/r2f_bench_$w.json 2> gpurun_out/r2f_bench_$w.err
done
timeout 400 python bench.py --workload e29 --steps 5 --warmup 3 > gpurun_out/r2f_bench_e29.json 2> gpurun_out/r2f_bench_e29.err
timeout 200 python bench.py --workload raw_wave --batch 64 --steps 10 --warmup 3 > gpurun_out/r2f_bench_raw_wave_b64.json 2> gpurun_out/r2f_bench_raw_wave_b64.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2f_bench_reference.json 2> gpurun_out/r2f_bench_reference.err
timeout 300 python tools/profile_step.py --warmup 3 --steps 2 --table > gpurun_out/r2f_table.log 2>&1
# launch list of one eager step (cold-cache, serialised: compare shares)
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2f_launches.csv python tools/profile_step.py --warmup 2 --steps 1 > gpurun_out/r2f_launches.log 2>&1
KEY="cpc_conv_dgrad b64 128x63x156->128x34x156 k30x1 s1x1 [tall_conv_tcgen05_128ch]"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tall128_conv_kernel -c 2 -o gpurun_out/r2f_dominant python tools/profile_kernel.py "$KEY" 1 > gpurun_out/r2f_ncu_dom.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:cqt_umma_kernel -c 1 -o gpurun_out/r2f_cqt python tools/profile_cqt.py > gpurun_out/r2f_ncu_cqt.log 2>&1
timeout 400 ncu --set full --clock-control none -k regex:"bn_bwd_reduce_kernel|bn_bwd_apply_kernel|bn_bwd_apply_packed_kernel|bn_stats_kernel|bn_apply_kernel|bn_apply_packed_kernel" -c 18 -o gpurun_out/r2f_bn python tools/profile_step.py --warmup 1 --steps 1 > gpurun_out/r2f_ncu_bn.log 2>&1
tail -3 gpurun_out/r2f_tests.log; tail -1 gpurun_out/r2f_smoke.log
for w in e24 raw_wave e20_bf16 long_context infonce_sweep e29 raw_wave_b64 reference; do python -c "
import json
try:
    d=json.load(open('gpurun_out/r2f_bench_$w.json')); print('$w', d.get('ms_per_step'), d.get('value'), d.get('e2e',{}).get('value'))
except Exception as e: print('$w', 'FAILED', e)"; done
