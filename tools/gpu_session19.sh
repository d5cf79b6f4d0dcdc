#!/bin/bash
# round-2 GPU session 19: class walks start at the class's cost level (skip keys): parity, timing, no-skip / no-split A/B
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 -x -k "extension" 2>&1 | tail -15) > gpurun_out/r02_pytest19.log 2>&1
tail -5 gpurun_out/r02_pytest19.log
timeout 900 python tools/auction_sweep.py 100000x1000000 ";" ";64" ";32" "256,2;" ";2048" > gpurun_out/r02_auction_sweep19.log 2>&1
cat gpurun_out/r02_auction_sweep19.log | cut -c1-260
(PM_TUNE_AUCTION=4 timeout 300 python tools/auction_scale.py 100000x1000000) > gpurun_out/r02_auction_trace19.log 2>&1
tail -1 gpurun_out/r02_auction_trace19.log
