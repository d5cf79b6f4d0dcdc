#!/bin/bash
# round-2 GPU session 12: proximity with the unlocated-tail partition; full suite; compute-sanitizer on the new kernels
mkdir -p gpurun_out
(time timeout 600 python -m pytest tests/test_gpu_proximity.py -m gpu -q --timeout 120 -x 2>&1 | tail -8) > gpurun_out/r02_prox7.log 2>&1
tail -4 gpurun_out/r02_prox7.log
(PM_TEST_BIG=1 timeout 900 python -m pytest tests/test_gpu_proximity.py -m gpu -q --timeout 300 -s -k "(100k and grid) or 1m" 2>&1 | tail -20) > gpurun_out/r02_prox_timing7.log 2>&1
tail -6 gpurun_out/r02_prox_timing7.log
(time timeout 1500 python -m pytest tests -m gpu -q --timeout 300 -x --ignore tests/test_gpu_proximity.py 2>&1 | tail -12) > gpurun_out/r02_pytest12.log 2>&1
tail -4 gpurun_out/r02_pytest12.log
(timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_proximity.py tests/test_gpu_sharded.py tests/test_gpu_variants.py tests/test_gpu_plugin_resident.py -m gpu -q --timeout 900 -x -k "not 100k and not 1m and not 100k_models" 2>&1 | tail -15) > gpurun_out/r02_sanitizer_memcheck.log 2>&1
tail -6 gpurun_out/r02_sanitizer_memcheck.log
(timeout 1200 compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_variants.py -m gpu -q --timeout 900 -x -k "ragged_shapes or kat_vectors or 200_models or more_options" 2>&1 | tail -15) > gpurun_out/r02_sanitizer_racecheck.log 2>&1
tail -6 gpurun_out/r02_sanitizer_racecheck.log
