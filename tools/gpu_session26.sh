#!/bin/bash
# round-2 GPU session 26 (final validation): whole suite, default bench line, auction trace + warm launch list, sanitizers on the new kernels
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -q --timeout 300 -x 2>&1 | tail -12) > gpurun_out/r02_pytest26.log 2>&1
tail -4 gpurun_out/r02_pytest26.log
(time timeout 600 python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err)
cut -c1-400 gpurun_out/r02_bench_final.json
(PM_TUNE_AUCTION=4 timeout 120 python tools/auction_scale.py 100000x1000000) > gpurun_out/r02_auction_trace26.log 2>&1
tail -1 gpurun_out/r02_auction_trace26.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --launch-skip 6000 -c 1200 --csv --log-file gpurun_out/auc_early26.csv python tools/auction_scale.py 100000x1000000 > gpurun_out/auc_early26.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --launch-skip 170000 -c 1200 --csv --log-file gpurun_out/auc_tail26.csv python tools/auction_scale.py 100000x1000000 > gpurun_out/auc_tail26.log 2>&1
(timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 280 -x -k "extension and (matches_self_oracle or eps_phases)" 2>&1 | tail -12) > gpurun_out/r02_sanitizer_memcheck_auction.log 2>&1
tail -4 gpurun_out/r02_sanitizer_memcheck_auction.log
(timeout 300 compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 280 -x -k "extension and matches_self_oracle" 2>&1 | tail -12) > gpurun_out/r02_sanitizer_racecheck_auction.log 2>&1
tail -4 gpurun_out/r02_sanitizer_racecheck_auction.log
