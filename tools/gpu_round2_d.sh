#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -s -k "cqt or scalogram_encoder or gradient_penalty or e24" > gpurun_out/r2k_cqt.log 2>&1
timeout 120 python tools/profile_cqt.py > gpurun_out/r2k_cqt_time.log 2>&1
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r2k_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2k_tests.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err
tail -3 gpurun_out/r2k_tests.log; cat gpurun_out/r2k_cqt_time.log
