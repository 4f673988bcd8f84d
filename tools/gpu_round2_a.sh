#!/bin/bash
# first GPU session of round 2: full GPU suite, the e24 golden under both filterbank paths, bench
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2a_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2a_tests.log
python -m pytest tests -m gpu -q -s -k "e24_full_size" > gpurun_out/r2a_e24_tensor.log 2>&1
CPC_NO_TENSOR_CQT=1 python -m pytest tests -m gpu -q -s -k "e24_full_size" > gpurun_out/r2a_e24_fp32cqt.log 2>&1
python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
tail -3 gpurun_out/r2a_tests.log
