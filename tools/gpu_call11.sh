mkdir -p gpurun_out/r2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bm25_fast_kernel -c 2 -o gpurun_out/r2/bm25_fast3 -f python tools/bm25_probe.py 1000000 1024 1 > gpurun_out/r2/ncu_bm25.log 2>&1; echo ncu $?; tail -2 gpurun_out/r2/ncu_bm25.log
