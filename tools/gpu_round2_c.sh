#!/bin/bash
mkdir -p gpurun_out
python tools/diag_e24_gates.py > gpurun_out/r2c_gates.log 2>&1
python tools/profile_cqt.py > gpurun_out/r2c_cqt_time.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:cqt_umma_kernel -c 1 -o gpurun_out/r2c_cqt python tools/profile_cqt.py > gpurun_out/r2c_ncu.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2c_cqt_launches.csv python tools/profile_cqt.py > /dev/null 2>&1
cat gpurun_out/r2c_gates.log gpurun_out/r2c_cqt_time.log
