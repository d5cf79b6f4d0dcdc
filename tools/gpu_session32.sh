#!/bin/bash
# round-2 GPU session 32: worker-major acceptance for wide model catalogues: variants parity, bench with 100k / 3000 model strings, default line
mkdir -p gpurun_out
(time timeout 300 python -m pytest tests/test_gpu_variants.py -m gpu -q --timeout 120 -x 2>&1 | tail -6) > gpurun_out/r02_pytest32.log 2>&1
tail -4 gpurun_out/r02_pytest32.log
for m in 100000 0; do
  timeout 200 python bench.py --models $m --steps 8 --warmup 3 --no-cpu --no-extras > gpurun_out/r02_bench_models32_$m.json 2> gpurun_out/r02_bench_models32_$m.err
  python -c "
import json;d=json.load(open('gpurun_out/r02_bench_models32_$m.json'));print($m, d['value'], d['kernel_ms_per_step'], d['roofline']['frac'], d['roofline']['other']['frac'])"
done
PM_TUNE_BUILD=7 timeout 200 python bench.py --models 100000 --steps 4 --warmup 3 --no-cpu --no-extras 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('global-memory table', d['value'], d['kernel_ms_per_step'])"
timeout 200 python bench.py --models 100000 --path fused-lean --steps 4 --warmup 3 --no-cpu --no-extras 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('fused-lean 100k models', d['value'], d['kernel_ms_per_step'])"
PM_TUNE_BUILD=7 timeout 200 python bench.py --models 100000 --path fused-lean --steps 4 --warmup 3 --no-cpu --no-extras 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('fused-lean 100k models, global-memory table', d['value'], d['kernel_ms_per_step'])"
