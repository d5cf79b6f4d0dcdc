#!/bin/bash
# round-2 GPU session 28: where the cluster-walk build spends its rounds: trace + warm launch lists (mid phase, tail)
mkdir -p gpurun_out
(PM_TUNE_AUCTION=4 timeout 100 python tools/auction_scale.py 100000x1000000) > gpurun_out/r02_auction_trace28.log 2>&1
tail -1 gpurun_out/r02_auction_trace28.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --launch-skip 30000 -c 900 --csv --log-file gpurun_out/auc_mid28.csv python tools/auction_scale.py 100000x1000000 > gpurun_out/auc_mid28.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --launch-skip 170000 -c 900 --csv --log-file gpurun_out/auc_tail28.csv python tools/auction_scale.py 100000x1000000 > gpurun_out/auc_tail28.log 2>&1
tail -1 gpurun_out/auc_tail28.log
