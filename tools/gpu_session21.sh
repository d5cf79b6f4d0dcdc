#!/bin/bash
# round-2 GPU session 21: warm-cache launch lists of auction rounds (ncu --cache-control none): kernel time vs wall time per round
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --launch-skip 7000 -c 1400 --csv --log-file gpurun_out/auc_early21.csv python tools/auction_scale.py 100000x1000000 > gpurun_out/auc_early21.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --launch-skip 200000 -c 1400 --csv --log-file gpurun_out/auc_tail21.csv python tools/auction_scale.py 100000x1000000 > gpurun_out/auc_tail21.log 2>&1
tail -1 gpurun_out/auc_tail21.log
