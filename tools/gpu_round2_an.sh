#!/bin/bash
# ReLU-epilogue gradient in one pass: AudioEncoder parity + raw_wave benches
mkdir -p gpurun_out
P=gpurun_out/r2j2
timeout 300 python -m pytest tests -m gpu -q -x -k "audio_encoder or raw or replay or conv" > ${P}_sub.log 2>&1; echo "subset rc=$?" >> ${P}_sub.log
tail -3 ${P}_sub.log
timeout 300 python bench.py --workload raw_wave --steps 10 --warmup 3 > ${P}_bench_raw_wave.json 2> ${P}_bench_raw_wave.err
timeout 300 python bench.py --workload raw_wave --batch 64 --steps 10 --warmup 3 > ${P}_bench_raw_wave_b64.json 2> ${P}_bench_raw_wave_b64.err
for w in raw_wave raw_wave_b64; do python -c "
import json
try:
    d=json.load(open('${P}_bench_$w.json')); print('$w', d.get('ms_per_step'), d.get('value'), d.get('e2e',{}).get('value'))
except Exception as e: print('$w', 'FAILED', e)"; done
