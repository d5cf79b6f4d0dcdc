#!/bin/bash
# round-2 GPU session 2: suite after the fast-predicate rewrite, cfg3 bench, ncu launch list + full capture of the two
# HBM kernels, proximity timings (100k both sweeps, 1M cooperative sweep) and an ncu capture of the cooperative sweep
mkdir -p gpurun_out
(time timeout 1200 python -m pytest tests -m gpu -q --timeout 300 -x --deselect tests/test_gpu_parity.py::test_cfg3_auction --ignore tests/test_gpu_proximity.py 2>&1 | tail -30) > gpurun_out/r02_pytest2.log 2>&1
tail -4 gpurun_out/r02_pytest2.log
(timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_cfg3_b.json) 2> gpurun_out/r02_bench_cfg3_b.err
cut -c1-300 gpurun_out/r02_bench_cfg3_b.json; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_cfg3_b.json')); print(d['roofline']['frac'], d['roofline']['other']['frac'], d['kernel_ms_per_step'], d['fused_lean']['ms_per_step'])"
(PM_TEST_BIG=1 timeout 900 python -m pytest tests/test_gpu_proximity.py -m gpu -q --timeout 300 -s -k "100k or 1m" 2>&1 | tail -20) > gpurun_out/r02_prox_timing.log 2>&1
tail -12 gpurun_out/r02_prox_timing.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-extras > gpurun_out/r02_ncu_a.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'pm_build_cost|pm_argmin' -s 40 -c 2 -o gpurun_out/r02_prof python bench.py --steps 1 --warmup 3 --no-cpu --no-extras > gpurun_out/r02_ncu_b.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'pm_proximity_grid' -c 1 -o gpurun_out/r02_prof_prox python -m pytest tests/test_gpu_proximity.py -m gpu -q -k "100k and grid" > gpurun_out/r02_ncu_c.log 2>&1
ls -la gpurun_out/*.ncu-rep; tail -3 gpurun_out/r02_ncu_b.log gpurun_out/r02_ncu_c.log
