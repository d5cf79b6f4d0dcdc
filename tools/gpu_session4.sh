#!/bin/bash
# round-2 GPU session 4: suite with the seed-parallel proximity phase and the resident-table plugin; proximity timings
mkdir -p gpurun_out
(time timeout 600 python -m pytest tests/test_gpu_proximity.py -m gpu -q --timeout 120 -x 2>&1 | tail -15) > gpurun_out/r02_prox2.log 2>&1
tail -5 gpurun_out/r02_prox2.log
(time timeout 1500 python -m pytest tests -m gpu -q --timeout 300 -x --deselect tests/test_gpu_parity.py::test_cfg3_auction --ignore tests/test_gpu_proximity.py 2>&1 | tail -30) > gpurun_out/r02_pytest4.log 2>&1
tail -5 gpurun_out/r02_pytest4.log
(PM_TEST_BIG=1 timeout 900 python -m pytest tests/test_gpu_proximity.py -m gpu -q --timeout 300 -s -k "100k or 1m" 2>&1 | tail -20) > gpurun_out/r02_prox_timing2.log 2>&1
tail -12 gpurun_out/r02_prox_timing2.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'pm_proximity_grid' -c 1 -o gpurun_out/r02_prof_prox2 python -m pytest tests/test_gpu_proximity.py -m gpu -q -k "100k and grid" > gpurun_out/r02_ncu_d.log 2>&1
tail -3 gpurun_out/r02_ncu_d.log
g++ -O2 -std=c++17 tools/host_bench.cpp -Iinclude -Lprotocol_b200 -lprime_match -lpthread -Wl,-rpath,$PWD/protocol_b200 -o /tmp/pm_host_bench 2>&1 | tail -3
(timeout 600 /tmp/pm_host_bench 1000000 2000) > gpurun_out/r02_host_bench_1m_b.txt 2>&1
head -9 gpurun_out/r02_host_bench_1m_b.txt
