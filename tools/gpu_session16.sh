#!/bin/bash
# round-2 GPU session 16: auction v7 (CTA bitonic selection; split walks stop at 33 good candidates): parity, time, launch lists
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 -x -k "extension" 2>&1 | tail -15) > gpurun_out/r02_pytest16.log 2>&1
tail -5 gpurun_out/r02_pytest16.log
(PM_TUNE_AUCTION=4 timeout 300 python tools/auction_scale.py 10000x100000 100000x1000000) > gpurun_out/r02_auction_trace16.log 2>&1
grep -v "auction batch" gpurun_out/r02_auction_trace16.log | tail -3
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 9000 -c 1800 --csv --log-file gpurun_out/auc_early16.csv python tools/auction_scale.py 100000x1000000 > gpurun_out/auc_early16.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 230000 -c 1800 --csv --log-file gpurun_out/auc_tail16.csv python tools/auction_scale.py 100000x1000000 > gpurun_out/auc_tail16.log 2>&1
tail -1 gpurun_out/auc_tail16.log
