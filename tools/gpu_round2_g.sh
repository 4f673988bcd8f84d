#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r2g_probe.log
for v in inline same_thread_global same_thread_local hook_local hook_global; do
  echo "=== $v" >> gpurun_out/r2g_probe.log
  timeout 90 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/probes/nccl_capture_probe.py $v >> gpurun_out/r2g_probe.log 2>&1
done
grep -a "VARIANT\|Error" gpurun_out/r2g_probe.log | head -40
