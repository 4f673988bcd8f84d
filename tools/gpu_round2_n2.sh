#!/bin/bash
# two GPUs of one box: NCCL parity test, e24 weak scaling N=1 vs N=2 on the same box, bf16 workload at N=2
mkdir -p gpurun_out
P=gpurun_out/r2h
timeout 300 python -m pytest tests -m gpu -q -k "nccl or rank" > ${P}_tests.log 2>&1; echo "tests rc=$?" >> ${P}_tests.log
tail -3 ${P}_tests.log
timeout 200 python bench.py --steps 20 --warmup 5 > ${P}_bench_n1.json 2> ${P}_bench_n1.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > ${P}_bench_n2.json 2> ${P}_bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload e20_bf16 --steps 10 --warmup 3 > ${P}_bench_e20_bf16_n2.json 2> ${P}_bench_e20_bf16_n2.err
for w in n1 n2 e20_bf16_n2; do python -c "
import json
try:
    d=json.loads(open('${P}_bench_$w.json').read().strip().splitlines()[-1]); print('$w', d.get('ms_per_step'), d.get('value'), d.get('e2e',{}).get('value'), d['clocks'])
except Exception as e: print('$w', 'FAILED', e)"; done
tail -2 ${P}_bench_n2.err
