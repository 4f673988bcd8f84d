#!/bin/bash
# one-plane front end in the bf16 operand mode: full suite, then (only when green) benches + the ncu captures that stamp
# profiles/ncu_traffic.json for the new sources
mkdir -p gpurun_out
P=gpurun_out/r2k2
timeout 600 python -m pytest tests -m gpu -q -s > ${P}_tests.log 2>&1; rc=$?; echo "tests rc=$rc" >> ${P}_tests.log
grep -a "e20 fp32\|cosine" ${P}_tests.log; tail -3 ${P}_tests.log
if [ $rc -ne 0 ]; then grep -a "Error\|assert\|FAILED" ${P}_tests.log | head -20; exit 1; fi
timeout 200 python tools/profile_step.py --warmup 3 --steps 2 --table > ${P}_table.log 2>&1; grep -a "wgrad.*tall_conv" ${P}_table.log | cut -c1-150
timeout 300 python bench.py --steps 20 --warmup 5 > ${P}_bench_e24.json 2> ${P}_bench_e24.err
KEY="cpc_conv_dgrad b64 128x63x156->128x34x156 k30x1 s1x1 [tall_conv_tcgen05_128ch]"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tall128_conv_kernel -c 2 -o ${P}_dominant python tools/profile_kernel.py "$KEY" 1 > ${P}_ncu_dom.log 2>&1
KEY2="cpc_conv_fwd b64 32x127x314->32x127x314 k64x1 s1x1 [tall_conv_tcgen05_32ch]"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tall_conv_kernel -c 2 -o ${P}_tall32 python tools/profile_kernel.py "$KEY2" 1 > ${P}_ncu_tall32.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:cqt_umma_kernel -c 1 -o ${P}_cqt python tools/profile_cqt.py > ${P}_ncu_cqt.log 2>&1
for w in e24; do python -c "
import json
try:
    d=json.load(open('${P}_bench_$w.json')); print('$w', d.get('ms_per_step'), d.get('value'), d.get('e2e',{}).get('value'), d['clocks']['sm_mhz'], d['metrics']['cqt_ms'])
except Exception as e: print('$w', 'FAILED', e)"; done
