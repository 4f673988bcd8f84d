#!/bin/bash
# usage: gpu_round2_multi.sh N [also_tests]   -- single-GPU line and N-GPU line of the e24 bench on the same box
N=$1
mkdir -p gpurun_out
if [ -n "$2" ]; then
  timeout 300 python -m pytest tests -m gpu -q -k "nccl" > gpurun_out/r2m_nccl_tests_n$N.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2m_nccl_tests_n$N.log
  tail -3 gpurun_out/r2m_nccl_tests_n$N.log
fi
timeout 200 python bench.py --steps 20 --warmup 5 > gpurun_out/r2m_bench_n1_box$N.json 2> gpurun_out/r2m_bench_n1_box$N.err
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2m_bench_n$N.json 2> gpurun_out/r2m_bench_n$N.err
for f in n1_box$N n$N; do python -c "
import json
try:
    d=json.load(open('gpurun_out/r2m_bench_$f.json')); print('$f', d.get('n_gpus'), d.get('ms_per_step'), d.get('value'), d.get('e2e',{}).get('value'), d.get('roofline',{}).get('traffic'))
except Exception as e: print('$f', 'FAILED', e)"; done
