#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -s -k "high_res_experiments or e29_default" > gpurun_out/r2y_e29.log 2>&1; echo "rc=$?" >> gpurun_out/r2y_e29.log
grep -a "e29\|e32\|passed\|failed\|Error\|^E " gpurun_out/r2y_e29.log | head -20
