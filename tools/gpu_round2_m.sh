#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r2m_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2m_tests.log
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r2m_tests2.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2m_tests2.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err
tail -3 gpurun_out/r2m_tests.log; tail -3 gpurun_out/r2m_tests2.log; grep -a "FAILED\|^E  " gpurun_out/r2m_tests*.log | head
python -c "
import json; d=json.load(open('gpurun_out/r2m_bench.json')); print(d['ms_per_step'], d['value'], d['metrics']['cqt_ms'], d['metrics']['cqt_hbm_gbs'])"
