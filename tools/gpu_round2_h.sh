#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 tests/nccl_parity_worker.py > gpurun_out/r2h_worker.log 2>&1
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2h_bench_n2.json 2> gpurun_out/r2h_bench_n2.err
grep -a "RANK_\|Error\|error" gpurun_out/r2h_worker.log | head -20; head -c 400 gpurun_out/r2h_bench_n2.json
