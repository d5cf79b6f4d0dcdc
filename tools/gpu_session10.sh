#!/bin/bash
# where does the 1M-node proximity sweep spend its time?  (source-level samples, few ncu passes)
mkdir -p gpurun_out
PM_TEST_BIG=1 timeout 1500 ncu --section SourceCounters --section WarpStateStats --section SchedulerStats --section MemoryWorkloadAnalysis --import-source on --clock-control none -k regex:'pm_proximity_grid' -c 1 -o gpurun_out/r02_prof_prox_1m python -m pytest tests/test_gpu_proximity.py -m gpu -q -k "1m" > gpurun_out/r02_ncu_i.log 2>&1
tail -3 gpurun_out/r02_ncu_i.log; ls -la gpurun_out/r02_prof_prox_1m.ncu-rep
