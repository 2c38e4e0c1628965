mkdir -p gpurun_out/r2
timeout 200 python tools/bm25_probe.py > gpurun_out/r2/bm25_probe.log 2>&1; echo probe $?; tail -2 gpurun_out/r2/bm25_probe.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bm25_ -c 8 -o gpurun_out/r2/bm25_fast -f python tools/bm25_probe.py 1000000 1024 1 > gpurun_out/r2/ncu_bm25.log 2>&1; echo ncu $?; tail -3 gpurun_out/r2/ncu_bm25.log
ls -la gpurun_out/r2/
