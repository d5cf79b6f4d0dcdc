#!/bin/bash
# 2-GPU bench rerun with per-rank logs
mkdir -p gpurun_out/tr_logs
N=$(nvidia-smi -L | wc -l)
timeout 900 python -u -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 --tee 3 --log-dir gpurun_out/tr_logs bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r02_bench_n${N}.json 2> gpurun_out/r02_bench_n${N}.err
echo "exit code $?"
tail -c 1500 gpurun_out/r02_bench_n${N}.err
find gpurun_out/tr_logs -type f | head; for f in $(find gpurun_out/tr_logs -name "*.log" -o -name "std*" | head -8); do echo "== $f"; tail -c 600 $f; done
head -c 600 gpurun_out/r02_bench_n${N}.json
