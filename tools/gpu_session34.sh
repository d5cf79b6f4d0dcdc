#!/bin/bash
# round-2 GPU session 34: ncu --set full of one early pool re-rank and one early class walk on the final auction kernels
mkdir -p gpurun_out
timeout 150 ncu --set full --import-source on --clock-control none --cache-control none --kernel-name regex:pm_auction_refill --launch-skip 1000 -c 1 -o gpurun_out/auc_refill34 -f python tools/auction_scale.py 100000x1000000 > gpurun_out/auc_refill34.log 2>&1
timeout 150 ncu --set full --import-source on --clock-control none --cache-control none --kernel-name regex:pm_auction_scan --launch-skip 2000 -c 1 -o gpurun_out/auc_scan34 -f python tools/auction_scale.py 100000x1000000 > gpurun_out/auc_scan34.log 2>&1
ls -la gpurun_out/*34.ncu-rep
