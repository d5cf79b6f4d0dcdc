#!/bin/bash
# round-2 GPU session 7: proximity after the MLP fix, rows-per-CTA A/B of the FastRow build kernel, full suite
mkdir -p gpurun_out
(time timeout 600 python -m pytest tests/test_gpu_proximity.py -m gpu -q --timeout 120 -x 2>&1 | tail -15) > gpurun_out/r02_prox4.log 2>&1
tail -4 gpurun_out/r02_prox4.log
(PM_TEST_BIG=1 timeout 900 python -m pytest tests/test_gpu_proximity.py -m gpu -q --timeout 300 -s -k "100k or 1m" 2>&1 | tail -20) > gpurun_out/r02_prox_timing4.log 2>&1
tail -10 gpurun_out/r02_prox_timing4.log
for tb in 5 0 6; do
  (PM_TUNE_BUILD=$tb timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-extras > gpurun_out/r02_bench_rows_tune${tb}.json) 2> gpurun_out/r02_bench_rows_tune${tb}.err
  python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_rows_tune${tb}.json').read().splitlines()[-1]); print('PM_TUNE_BUILD=${tb}', d['value'], d['roofline']['kernel'], d['roofline']['frac'], d['roofline']['other']['frac'], d['kernel_ms_per_step'])"
done
(time timeout 1500 python -m pytest tests -m gpu -q --timeout 300 -x --deselect tests/test_gpu_parity.py::test_cfg3_auction --ignore tests/test_gpu_proximity.py 2>&1 | tail -30) > gpurun_out/r02_pytest7.log 2>&1
tail -4 gpurun_out/r02_pytest7.log
g++ -O2 -std=c++17 tools/host_bench.cpp -Iinclude -Lprotocol_b200 -lprime_match -lpthread -Wl,-rpath,$PWD/protocol_b200 -o /tmp/pm_host_bench 2>&1 | tail -3
(timeout 600 /tmp/pm_host_bench 1000000 2000) > gpurun_out/r02_host_bench_1m_d.txt 2>&1
head -9 gpurun_out/r02_host_bench_1m_d.txt
