#!/bin/bash
# round-2 GPU session 8: staggered proximity scans; final single-GPU captures (bench, launch list, ncu --set full)
mkdir -p gpurun_out
(time timeout 600 python -m pytest tests/test_gpu_proximity.py -m gpu -q --timeout 120 -x 2>&1 | tail -15) > gpurun_out/r02_prox5.log 2>&1
tail -4 gpurun_out/r02_prox5.log
(PM_TEST_BIG=1 timeout 900 python -m pytest tests/test_gpu_proximity.py -m gpu -q --timeout 300 -s -k "100k or 1m" 2>&1 | tail -20) > gpurun_out/r02_prox_timing5.log 2>&1
tail -10 gpurun_out/r02_prox_timing5.log
(time timeout 1500 python -m pytest tests -m gpu -q --timeout 300 -x --deselect tests/test_gpu_parity.py::test_cfg3_auction --ignore tests/test_gpu_proximity.py 2>&1 | tail -30) > gpurun_out/r02_pytest8.log 2>&1
tail -4 gpurun_out/r02_pytest8.log
(timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_cfg3_final.json) 2> gpurun_out/r02_bench_cfg3_final.err
python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_cfg3_final.json').read().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['roofline']['kernel'], d['roofline']['frac'], d['roofline']['other']['frac'], d['kernel_ms_per_step'], d['cpu_baseline']['value'], d['fused_lean']['ms_per_step'], d['auction'])"
(timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_reference_cfg3.json) 2> gpurun_out/r02_bench_reference_cfg3.err
cut -c1-300 gpurun_out/r02_bench_reference_cfg3.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_final.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-extras > gpurun_out/r02_ncu_f.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'pm_build_cost|pm_argmin' -s 40 -c 2 -o gpurun_out/r02_prof_final python bench.py --steps 1 --warmup 3 --no-cpu --no-extras > gpurun_out/r02_ncu_g.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'pm_proximity_grid' -c 1 -o gpurun_out/r02_prof_prox_final python -m pytest tests/test_gpu_proximity.py -m gpu -q -k "100k and grid" > gpurun_out/r02_ncu_h.log 2>&1
ls -la gpurun_out/*final*.ncu-rep
(timeout 300 python tools/sass_evidence.py > gpurun_out/r02_sass_evidence.txt) 2>&1 | tail -2
