#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r2l_rounds.log
for r in 8 16 32 64 255; do
  echo "=== CPC_CQT_ROUND=$r" >> gpurun_out/r2l_rounds.log
  CPC_CQT_ROUND=$r timeout 120 python tools/profile_cqt.py >> gpurun_out/r2l_rounds.log 2>&1
  CPC_CQT_ROUND=$r timeout 200 python -m pytest tests -m gpu -q -s -k "fp32_exact" 2>&1 | grep -a "K=16384\|K= 8192\|K= 1024\|passed\|failed" >> gpurun_out/r2l_rounds.log
done
cat gpurun_out/r2l_rounds.log
