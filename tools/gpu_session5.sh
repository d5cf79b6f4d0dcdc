#!/bin/bash
# round-2 GPU session 5 (2+ GPUs): the sharded pass inside the library on real peers — pm_multi (one process) in the
# test-suite, pm_comm (one process per GPU) through bench.py under torchrun
mkdir -p gpurun_out
nvidia-smi -L
(time timeout 600 python -m pytest tests/test_gpu_sharded.py -m gpu -q --timeout 300 -x 2>&1 | tail -15) > gpurun_out/r02_sharded.log 2>&1
tail -6 gpurun_out/r02_sharded.log
N=$(nvidia-smi -L | wc -l)
(time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r02_bench_n${N}.json) 2> gpurun_out/r02_bench_n${N}.err
tail -5 gpurun_out/r02_bench_n${N}.err
python - <<PY
import json
d=json.load(open('gpurun_out/r02_bench_n${N}.json'))
print({k:d[k] for k in ('value','n_gpus','ms_per_step','ranks_agree','groups_formed','kernel_ms_per_step')}); print(d['config']['workload']); print(d['e2e']); print(d['roofline']['frac'], d['roofline']['other']['frac'])
PY
(time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_ref_n${N}.json) 2> gpurun_out/r02_bench_ref_n${N}.err
cut -c1-400 gpurun_out/r02_bench_ref_n${N}.json
