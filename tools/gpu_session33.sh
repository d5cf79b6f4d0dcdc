#!/bin/bash
# round-2 GPU session 33 (last): whole suite + smoke on the final library
mkdir -p gpurun_out
(time timeout 400 python -m pytest tests -m gpu -q --timeout 200 -x 2>&1 | tail -8) > gpurun_out/r02_pytest33.log 2>&1
tail -4 gpurun_out/r02_pytest33.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
