#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r2u_nccl_ctas.log
for c in default 2 4 8; do
  if [ "$c" = "default" ]; then unset NCCL_MAX_CTAS; else export NCCL_MAX_CTAS=$c; fi
  timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29530 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/r2u_n2_$c.json 2> gpurun_out/r2u_n2_$c.err
  python -c "
import json
d=json.load(open('gpurun_out/r2u_n2_$c.json')); print('NCCL_MAX_CTAS=$c', round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3))" >> gpurun_out/r2u_nccl_ctas.log 2>&1
done
unset NCCL_MAX_CTAS
timeout 150 python bench.py --steps 30 --warmup 5 > gpurun_out/r2u_n1.json 2> gpurun_out/r2u_n1.err
python -c "
import json
d=json.load(open('gpurun_out/r2u_n1.json')); print('N=1', round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3))" >> gpurun_out/r2u_nccl_ctas.log 2>&1
cat gpurun_out/r2u_nccl_ctas.log
