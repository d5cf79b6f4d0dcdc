#!/bin/bash
# round-2 GPU session 18: class walks only as deep as the waiting asks need: parity, timing against the 33-deep walks, pool knobs again
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 -x -k "extension" 2>&1 | tail -15) > gpurun_out/r02_pytest18.log 2>&1
tail -5 gpurun_out/r02_pytest18.log
timeout 900 python tools/auction_sweep.py 100000x1000000 ";" ";128" "256,2;" "64,1;" ";32" ";160" > gpurun_out/r02_auction_sweep18.log 2>&1
cat gpurun_out/r02_auction_sweep18.log | cut -c1-260
(PM_TUNE_AUCTION=4 timeout 300 python tools/auction_scale.py 100000x1000000) > gpurun_out/r02_auction_trace18.log 2>&1
tail -1 gpurun_out/r02_auction_trace18.log
