mkdir -p gpurun_out/r2
timeout 300 python -m pytest tests/test_gpu_bm25_rrf.py -x -q -m gpu > gpurun_out/r2/t_bm25.log 2>&1; echo bm25 $?; tail -5 gpurun_out/r2/t_bm25.log
timeout 120 python tools/peak_probe.py gpurun_out/r2/peaks.json > gpurun_out/r2/peaks.log 2>&1; echo probe $?; tail -3 gpurun_out/r2/peaks.log
timeout 120 python __graft_entry__.py smoke > gpurun_out/r2/smoke.log 2>&1; echo smoke $?; tail -3 gpurun_out/r2/smoke.log
timeout 420 python -m pytest tests/test_gpu_at_size.py -x -q -m gpu --durations=8 > gpurun_out/r2/t_size.log 2>&1; echo size $?; tail -25 gpurun_out/r2/t_size.log
