#!/bin/bash
mkdir -p gpurun_out
timeout 250 python -m pytest tests -m gpu -q -s -k "two_rank" > gpurun_out/r2n_nccl.log 2>&1
grep -a "RANK_\|passed\|failed" gpurun_out/r2n_nccl.log | head -20
