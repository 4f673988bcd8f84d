#!/bin/bash
# final evidence of round 2 (second session): full GPU suite, smoke, every workload's bench line, the reference arm,
# launch list, ncu --set full of the dominant kernel / the CQT / the rebalanced 64x1 row-streaming kernel
mkdir -p gpurun_out
P=gpurun_out/r2g
timeout 900 python -m pytest tests -m gpu -q > ${P}_tests.log 2>&1; echo "tests rc=$?" >> ${P}_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > ${P}_smoke.log 2>&1
timeout 300 python bench.py --steps 20 --warmup 5 > ${P}_bench_e24.json 2> ${P}_bench_e24.err
for w in raw_wave e20_bf16 long_context infonce_sweep; do
  timeout 400 python bench.py --workload $w --steps 10 --warmup 3 > ${P}_bench_$w.json 2> ${P}_bench_$w.err
done
timeout 400 python bench.py --workload e29 --steps 5 --warmup 3 > ${P}_bench_e29.json 2> ${P}_bench_e29.err
timeout 200 python bench.py --workload raw_wave --batch 64 --steps 10 --warmup 3 > ${P}_bench_raw_wave_b64.json 2> ${P}_bench_raw_wave_b64.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > ${P}_bench_reference.json 2> ${P}_bench_reference.err
timeout 300 python tools/profile_step.py --warmup 3 --steps 2 --table > ${P}_table.log 2>&1
# launch list of eager steps (cold-cache, serialised: compare shares)
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file ${P}_launches.csv python tools/profile_step.py --warmup 2 --steps 1 > ${P}_launches.log 2>&1
KEY="cpc_conv_dgrad b64 128x63x156->128x34x156 k30x1 s1x1 [tall_conv_tcgen05_128ch]"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tall128_conv_kernel -c 2 -o ${P}_dominant python tools/profile_kernel.py "$KEY" 1 > ${P}_ncu_dom.log 2>&1
KEY2="cpc_conv_fwd b64 32x127x314->32x127x314 k64x1 s1x1 [tall_conv_tcgen05_32ch]"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tall_conv_kernel -c 2 -o ${P}_tall32 python tools/profile_kernel.py "$KEY2" 1 > ${P}_ncu_tall32.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:cqt_umma_kernel -c 1 -o ${P}_cqt python tools/profile_cqt.py > ${P}_ncu_cqt.log 2>&1
tail -3 ${P}_tests.log; tail -1 ${P}_smoke.log
for w in e24 raw_wave e20_bf16 long_context infonce_sweep e29 raw_wave_b64 reference; do python -c "
import json
try:
    d=json.load(open('${P}_bench_$w.json')); print('$w', d.get('ms_per_step'), d.get('value'), d.get('e2e',{}).get('value'))
except Exception as e: print('$w', 'FAILED', e)"; done
ls -la gpurun_out/r2g_* | awk '{print $5, $9}'
