#!/bin/bash
# round-2 GPU session 20: walk items drawn from a counter, one class per CTA from the first round on, refill on 8 CTAs per SM
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 -x -k "extension" 2>&1 | tail -15) > gpurun_out/r02_pytest20.log 2>&1
tail -5 gpurun_out/r02_pytest20.log
(PM_TUNE_AUCTION=4 timeout 300 python tools/auction_scale.py 10000x100000 100000x1000000) > gpurun_out/r02_auction_trace20.log 2>&1
grep -v "auction batch" gpurun_out/r02_auction_trace20.log | tail -2
grep "rounds=32 " gpurun_out/r02_auction_trace20.log
(PM_TUNE_AUCTION=4100 timeout 300 python tools/auction_scale.py 100000x1000000) > gpurun_out/r02_auction_trace20_tpc8.log 2>&1
grep -v "auction batch" gpurun_out/r02_auction_trace20_tpc8.log | tail -1
grep "rounds=32 " gpurun_out/r02_auction_trace20_tpc8.log
