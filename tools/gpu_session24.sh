#!/bin/bash
# round-2 GPU session 24: selections sort packed 64-bit keys when every reachable cost fits: parity, time
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 -x -k "extension" 2>&1 | tail -15) > gpurun_out/r02_pytest24.log 2>&1
tail -5 gpurun_out/r02_pytest24.log
(timeout 300 python tools/auction_scale.py 10000x100000 100000x1000000) > gpurun_out/r02_auction_scale24.log 2>&1
cat gpurun_out/r02_auction_scale24.log | cut -c1-250
