mkdir -p gpurun_out/r2
timeout 300 python -m pytest tests/test_gpu_bm25_rrf.py tests/test_gpu_at_size.py -x -q -m gpu -k "bm25 or config3" > gpurun_out/r2/t_bm25.log 2>&1; echo tests $?; tail -3 gpurun_out/r2/t_bm25.log
timeout 200 python tools/bm25_probe.py > gpurun_out/r2/bm25_probe.log 2>&1; tail -1 gpurun_out/r2/bm25_probe.log
