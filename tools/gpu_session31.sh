#!/bin/bash
# round-2 GPU session 31: the default bench workload with wide model catalogues (the kernel variants a permissionless pool selects)
mkdir -p gpurun_out
for m in 200 3000 100000; do
  timeout 200 python bench.py --models $m --steps 8 --warmup 3 --no-cpu --no-extras > gpurun_out/r02_bench_models_$m.json 2> gpurun_out/r02_bench_models_$m.err
  python -c "
import json;d=json.load(open('gpurun_out/r02_bench_models_$m.json'));print($m, d['value'], d['kernel_ms_per_step'], d['roofline']['frac'], d['roofline']['other']['frac'], d['detail'])"
done
