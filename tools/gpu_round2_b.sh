#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s -k "cqt or scalogram_encoder or gradient_penalty or training_steps or e24" > gpurun_out/r2b_cqt.log 2>&1
python -m pytest tests -m gpu -q > gpurun_out/r2b_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2b_tests.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
tail -3 gpurun_out/r2b_tests.log
