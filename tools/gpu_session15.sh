#!/bin/bash
# round-2 GPU session 15: auction v7 (CTA bitonic selection, complete lists for scarce classes): parity, then time
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 -x -k "extension" 2>&1 | tail -15) > gpurun_out/r02_pytest15.log 2>&1
tail -5 gpurun_out/r02_pytest15.log
(PM_TUNE_AUCTION=4 timeout 300 python tools/auction_scale.py 10000x100000 100000x1000000) > gpurun_out/r02_auction_trace15.log 2>&1
grep -v "auction batch" gpurun_out/r02_auction_trace15.log | tail -3
(PM_TUNE_AUCTION=64 timeout 300 python tools/auction_scale.py 100000x1000000) > gpurun_out/r02_auction_nolists15.log 2>&1
tail -1 gpurun_out/r02_auction_nolists15.log
