#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s -k "cqt or scalogram_encoder or gradient_penalty or e24" > gpurun_out/r2e_cqt.log 2>&1
python tools/profile_cqt.py > gpurun_out/r2e_cqt_time.log 2>&1
python -m pytest tests -m gpu -q > gpurun_out/r2e_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2e_tests.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err
tail -3 gpurun_out/r2e_tests.log; cat gpurun_out/r2e_cqt_time.log
