#!/bin/bash
# 8-GPU box: the driver's scaling launches (N = 8, 4), each under a timeout
mkdir -p gpurun_out
for n in 8 4; do
timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r2t_bench_n$n.json 2> gpurun_out/r2t_bench_n$n.err
echo "n=$n rc=$?"
done
python -c "
import json
for n in (8,4):
    try:
        d=json.load(open('gpurun_out/r2t_bench_n%d.json'%n)); print(n, d['ms_per_step'], d['value'], d['e2e']['value'], d.get('overlap','')[:60])
    except Exception as e: print(n, 'FAILED', e)"
