#!/bin/bash
# round-2 GPU session 25: packed claim (the atomicMax on (bid, bidder) is the claim): parity, time
mkdir -p gpurun_out
(time timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 120 -x -k "extension" 2>&1 | tail -15) > gpurun_out/r02_pytest25.log 2>&1
tail -5 gpurun_out/r02_pytest25.log
(timeout 120 python tools/auction_scale.py 10000x100000 100000x1000000) > gpurun_out/r02_auction_scale25.log 2>&1
cat gpurun_out/r02_auction_scale25.log | cut -c1-250
