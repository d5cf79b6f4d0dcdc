#!/bin/bash
# round-2 GPU session 22: auction after the TPC=8 clean-up: parity, time; ncu --set full of one early refill and one early walk
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 -x -k "extension" 2>&1 | tail -15) > gpurun_out/r02_pytest22.log 2>&1
tail -5 gpurun_out/r02_pytest22.log
(timeout 300 python tools/auction_scale.py 10000x100000 100000x1000000) > gpurun_out/r02_auction_scale22.log 2>&1
cat gpurun_out/r02_auction_scale22.log | cut -c1-250
timeout 600 ncu --set full --import-source on --clock-control none --cache-control none --kernel-name regex:pm_auction_refill --launch-skip 1000 -c 1 -o gpurun_out/auc_refill22 -f python tools/auction_scale.py 100000x1000000 > gpurun_out/auc_refill22.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none --cache-control none --kernel-name regex:pm_auction_scan --launch-skip 2000 -c 1 -o gpurun_out/auc_scan22 -f python tools/auction_scale.py 100000x1000000 > gpurun_out/auc_scan22.log 2>&1
ls -la gpurun_out/*.ncu-rep
