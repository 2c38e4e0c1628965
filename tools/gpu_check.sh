#!/bin/bash
# Quick end-of-round check on one GPU: the GPU test suite and smoke() with the library as built.
O=gpurun_out/r2/final; mkdir -p $O
timeout -k 10 1200 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest $?"; tail -3 $O/pytest_gpu.log | cut -c1-300
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; echo "smoke $?"; tail -2 $O/smoke.log | cut -c1-300
