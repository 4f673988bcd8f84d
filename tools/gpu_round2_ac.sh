#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/profile_step.py --warmup 3 --steps 2 --table > gpurun_out/r2ac_table.log 2>&1
tail -3 gpurun_out/r2ac_table.log
