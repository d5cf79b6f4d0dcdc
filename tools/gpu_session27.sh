#!/bin/bash
# round-2 GPU session 27: class walks shared by the CTAs of a thread-block cluster (DSMEM counts and lists): parity, time
mkdir -p gpurun_out
(time timeout 240 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 100 -x -k "extension" 2>&1 | tail -15) > gpurun_out/r02_pytest27.log 2>&1
tail -5 gpurun_out/r02_pytest27.log
(timeout 100 python tools/auction_scale.py 10000x100000 100000x1000000) > gpurun_out/r02_auction_scale27.log 2>&1
cat gpurun_out/r02_auction_scale27.log | cut -c1-250
(PM_TUNE_AUCTION=32 timeout 100 python tools/auction_scale.py 100000x1000000) > gpurun_out/r02_auction_scale27_nocluster.log 2>&1
cat gpurun_out/r02_auction_scale27_nocluster.log | cut -c1-250
