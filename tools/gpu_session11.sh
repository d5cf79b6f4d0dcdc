#!/bin/bash
# round-2 GPU session 11: proximity after the code-size fix (parity + timings)
mkdir -p gpurun_out
(time timeout 600 python -m pytest tests/test_gpu_proximity.py tests/test_gpu_parity.py -m gpu -q --timeout 120 -x -k "proximity or prox or mixed_group or larger_than or hand_down or pairs" 2>&1 | tail -15) > gpurun_out/r02_prox6.log 2>&1
tail -4 gpurun_out/r02_prox6.log
(PM_TEST_BIG=1 timeout 900 python -m pytest tests/test_gpu_proximity.py -m gpu -q --timeout 300 -s -k "(100k and grid) or 1m" 2>&1 | tail -20) > gpurun_out/r02_prox_timing6.log 2>&1
tail -8 gpurun_out/r02_prox_timing6.log
