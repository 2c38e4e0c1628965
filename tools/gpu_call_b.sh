#!/bin/bash
mkdir -p gpurun_out/r2
timeout -k 10 100 python tools/tc_diag.py 4096 128 16 8 2>&1 | tail -12 | cut -c1-400
timeout -k 10 100 python tools/tc_diag.py 100000 768 130 50 2>&1 | tail -8 | cut -c1-300
timeout -k 10 400 python -m pytest tests/test_gpu_tc.py tests/test_gpu_hamming.py -x -q 2>&1 | tail -8 | cut -c1-300
timeout -k 10 100 python tools/tc_probe.py 1000000 768 1024 400 2>&1 | grep ms_per_call
timeout -k 10 100 python tools/tc_probe.py 1000000 768 256 200 2>&1 | grep ms_per_call
timeout -k 10 100 python tools/tc_probe.py 12500000 1024 256 40 2>&1 | grep ms_per_call
