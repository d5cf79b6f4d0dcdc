#!/bin/bash
# round-2 GPU session 30 (2 GPUs): the sharded tests and the multi-GPU bench line on the final library
mkdir -p gpurun_out
(time timeout 400 python -m pytest tests/test_gpu_sharded.py -m gpu -q --timeout 200 -x 2>&1 | tail -6) > gpurun_out/r02_sharded_2gpu_final.log 2>&1
tail -4 gpurun_out/r02_sharded_2gpu_final.log
(time timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02_bench_n2_final.json 2> gpurun_out/r02_bench_n2_final.err)
cut -c1-900 gpurun_out/r02_bench_n2_final.json; tail -3 gpurun_out/r02_bench_n2_final.err
