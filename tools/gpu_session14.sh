#!/bin/bash
# round-2 GPU session 14: reputation parity; where an auction round's time goes (ncu launch lists of two windows of rounds)
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 -x -k "extension" 2>&1 | tail -8) > gpurun_out/r02_pytest14.log 2>&1
tail -4 gpurun_out/r02_pytest14.log
(PM_TUNE_AUCTION=4 timeout 300 python tools/auction_scale.py 100000x1000000) > gpurun_out/r02_auction_trace14.log 2>&1
tail -2 gpurun_out/r02_auction_trace14.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 9000 -c 1800 --csv --log-file gpurun_out/auc_early.csv python tools/auction_scale.py 100000x1000000 > gpurun_out/auc_early.log 2>&1
tail -2 gpurun_out/auc_early.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 230000 -c 1800 --csv --log-file gpurun_out/auc_tail.csv python tools/auction_scale.py 100000x1000000 > gpurun_out/auc_tail.log 2>&1
tail -2 gpurun_out/auc_tail.log
