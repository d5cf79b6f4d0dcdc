#!/bin/bash
# round-2 GPU session 3: the new plugin / anchor tests and the host-side rows at 1M nodes (tools/host_bench.cpp)
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_gpu_plugin_resident.py tests/test_gpu_plugin.py tests/test_gpu_parity.py -m gpu -q --timeout 300 -x -k "resident or delta or merge or plugin or scipy or group" 2>&1 | tail -15) > gpurun_out/r02_pytest3.log 2>&1
tail -6 gpurun_out/r02_pytest3.log
g++ -O2 -std=c++17 tools/host_bench.cpp -Iinclude -Lprotocol_b200 -lprime_match -lpthread -Wl,-rpath,$PWD/protocol_b200 -o /tmp/pm_host_bench 2>&1 | tail -3
(timeout 600 /tmp/pm_host_bench 1000000 2000) > gpurun_out/r02_host_bench_1m.txt 2>&1
cat gpurun_out/r02_host_bench_1m.txt
(timeout 300 /tmp/pm_host_bench 100000 2000) > gpurun_out/r02_host_bench_100k.txt 2>&1
cat gpurun_out/r02_host_bench_100k.txt
