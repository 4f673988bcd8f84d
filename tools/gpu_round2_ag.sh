#!/bin/bash
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"bn_bwd_reduce_kernel|bn_bwd_apply_kernel|bn_bwd_apply_packed_kernel|bn_stats_kernel|bn_apply_kernel" -c 16 -o gpurun_out/r2ag_bn python tools/profile_step.py --warmup 1 --steps 1 > gpurun_out/r2ag_ncu.log 2>&1
tail -3 gpurun_out/r2ag_ncu.log
ls -la gpurun_out/r2ag_bn.ncu-rep
