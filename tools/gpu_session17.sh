#!/bin/bash
# round-2 GPU session 17: auction with apply+compact+advance in one kernel: parity, then a sweep of the pool-fill / re-sort knobs
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 -x -k "extension" 2>&1 | tail -15) > gpurun_out/r02_pytest17.log 2>&1
tail -5 gpurun_out/r02_pytest17.log
timeout 900 python tools/auction_sweep.py 100000x1000000 ";" "256,4;" "256,2;" "128,4;" "64,1;" "512,8;" "256,16;" ";512" ";2048" "256,4;512" > gpurun_out/r02_auction_sweep17.log 2>&1
cat gpurun_out/r02_auction_sweep17.log | cut -c1-260
