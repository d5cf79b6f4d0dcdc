#!/bin/bash
# round-2 GPU session 1: the whole -m gpu suite, the new proximity sweep, the experimental banded sweep, a first cfg3 bench line
mkdir -p gpurun_out
nvidia-smi -L; nproc; cat /sys/fs/cgroup/cpu.max
(time timeout 1200 python -m pytest tests -m gpu -q --timeout 300 -x --deselect tests/test_gpu_parity.py::test_cfg3_auction --ignore tests/test_gpu_proximity.py 2>&1 | tail -40) > gpurun_out/r02_pytest1.log 2>&1
(time timeout 900 python -m pytest tests/test_gpu_proximity.py -m gpu -q --timeout 200 -s 2>&1 | tail -40) > gpurun_out/r02_prox.log 2>&1
(PM_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "banded or experimental" 2>&1 | tail -30) > gpurun_out/r02_banded.log 2>&1
(timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_cfg3_a.json) 2> gpurun_out/r02_bench_cfg3_a.err
tail -8 gpurun_out/r02_pytest1.log; tail -12 gpurun_out/r02_prox.log; tail -8 gpurun_out/r02_banded.log; cut -c1-900 gpurun_out/r02_bench_cfg3_a.json; tail -5 gpurun_out/r02_bench_cfg3_a.err
