#!/bin/bash
# round-2 GPU session 9 (8 GPUs): BASELINE configs[4] — 1M asks x 8M workers over 8 GPUs — through bench.py under
# torchrun (pm_comm inside the library), and pm_multi (one process, 8 devices) in the test-suite
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l); echo "GPUs: $N"
(time timeout 600 python -m pytest tests/test_gpu_sharded.py -m gpu -q --timeout 300 -x -k multi 2>&1 | tail -8) > gpurun_out/r02_sharded_n${N}.log 2>&1
tail -4 gpurun_out/r02_sharded_n${N}.log
(time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r02_bench_n${N}.json) 2> gpurun_out/r02_bench_n${N}.err
echo "bench rc $?"; tail -c 600 gpurun_out/r02_bench_n${N}.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_n${N}.json').read().splitlines()[-1])
print({k:d.get(k) for k in ('value','n_gpus','ms_per_step','ranks_agree','groups_formed','kernel_ms_per_step','clocks')}); print(d['config']['workload']); print(d['e2e']); print(d['roofline']['kernel'], d['roofline']['frac'], d['roofline']['other']['frac'])
PY
