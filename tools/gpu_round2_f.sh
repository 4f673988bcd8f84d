#!/bin/bash
# 2-GPU session: NCCL parity test, bench at N=2 (graph with captured overlapped all-reduce); every step under `timeout`
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q -s -k "two_rank" > gpurun_out/r2f_nccl.log 2>&1
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2f_bench_n2.json 2> gpurun_out/r2f_bench_n2.err
timeout 240 python bench.py --steps 10 --warmup 3 > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_bench_n1.err
tail -5 gpurun_out/r2f_nccl.log; tail -3 gpurun_out/r2f_bench_n2.err
