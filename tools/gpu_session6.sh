#!/bin/bash
# round-2 GPU session 6: pruned seed-parallel proximity phase; FastRow build kernel A/B against the staged-CSR form
mkdir -p gpurun_out
(time timeout 600 python -m pytest tests/test_gpu_proximity.py -m gpu -q --timeout 120 -x 2>&1 | tail -15) > gpurun_out/r02_prox3.log 2>&1
tail -4 gpurun_out/r02_prox3.log
(time timeout 1500 python -m pytest tests -m gpu -q --timeout 300 -x --deselect tests/test_gpu_parity.py::test_cfg3_auction --ignore tests/test_gpu_proximity.py 2>&1 | tail -30) > gpurun_out/r02_pytest6.log 2>&1
tail -4 gpurun_out/r02_pytest6.log
(PM_TEST_BIG=1 timeout 900 python -m pytest tests/test_gpu_proximity.py -m gpu -q --timeout 300 -s -k "100k or 1m" 2>&1 | tail -20) > gpurun_out/r02_prox_timing3.log 2>&1
tail -10 gpurun_out/r02_prox_timing3.log
for tb in 1 0; do
  (PM_TUNE_BUILD=$tb timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-extras > gpurun_out/r02_bench_build_tune${tb}.json) 2> gpurun_out/r02_bench_build_tune${tb}.err
  python -c "
import json; d=json.load(open('gpurun_out/r02_bench_build_tune${tb}.json')); print('PM_TUNE_BUILD=${tb}', d['value'], d['roofline']['kernel'], d['roofline']['frac'], d['roofline']['other']['frac'], d['kernel_ms_per_step'])"
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'pm_build_cost' -s 20 -c 1 -o gpurun_out/r02_prof_build2 python bench.py --steps 1 --warmup 3 --no-cpu --no-extras > gpurun_out/r02_ncu_e.log 2>&1
tail -2 gpurun_out/r02_ncu_e.log
g++ -O2 -std=c++17 tools/host_bench.cpp -Iinclude -Lprotocol_b200 -lprime_match -lpthread -Wl,-rpath,$PWD/protocol_b200 -o /tmp/pm_host_bench 2>&1 | tail -3
(timeout 600 /tmp/pm_host_bench 1000000 2000) > gpurun_out/r02_host_bench_1m_c.txt 2>&1
head -9 gpurun_out/r02_host_bench_1m_c.txt
