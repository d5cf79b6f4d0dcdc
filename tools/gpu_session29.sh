#!/bin/bash
# round-2 GPU session 29 (final): whole suite on the final library, smoke(), auction time, default bench line
mkdir -p gpurun_out
(time timeout 600 python -m pytest tests -m gpu -q --timeout 200 -x 2>&1 | tail -8) > gpurun_out/r02_pytest29.log 2>&1
tail -4 gpurun_out/r02_pytest29.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
(timeout 100 python tools/auction_scale.py 10000x100000 100000x1000000) 2>&1 | cut -c1-250
(time timeout 400 python bench.py > gpurun_out/r02_bench_final2.json 2> gpurun_out/r02_bench_final2.err)
cut -c1-300 gpurun_out/r02_bench_final2.json
