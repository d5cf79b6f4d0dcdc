#!/bin/bash
# round-2 GPU session 13 (re-entry): validate the unlocated-tail partition commit; full suite; default bench line
mkdir -p gpurun_out
(time timeout 600 python -m pytest tests/test_gpu_proximity.py -m gpu -q --timeout 120 -x 2>&1 | tail -8) > gpurun_out/r02_prox7.log 2>&1
tail -4 gpurun_out/r02_prox7.log
(PM_TEST_BIG=1 timeout 900 python -m pytest tests/test_gpu_proximity.py -m gpu -q --timeout 300 -s -k "(100k and grid) or 1m" 2>&1 | tail -20) > gpurun_out/r02_prox_timing7.log 2>&1
tail -6 gpurun_out/r02_prox_timing7.log
(time timeout 1500 python -m pytest tests -m gpu -q --timeout 300 -x --ignore tests/test_gpu_proximity.py 2>&1 | tail -12) > gpurun_out/r02_pytest13.log 2>&1
tail -4 gpurun_out/r02_pytest13.log
(time timeout 600 python bench.py > gpurun_out/r02_bench_default13.json 2> gpurun_out/r02_bench_default13.err)
cat gpurun_out/r02_bench_default13.json | cut -c1-1500
