mkdir -p gpurun_out/r2
timeout 120 python tools/peak_probe.py gpurun_out/r2/peaks.json > gpurun_out/r2/peaks.log 2>&1; echo probe $?; tail -2 gpurun_out/r2/peaks.log
cp gpurun_out/r2/peaks.json profiles/r2_peaks.json
timeout 600 python -m pytest tests -x -q -m gpu --durations=10 > gpurun_out/r2/t_all.log 2>&1; echo tests $?; tail -25 gpurun_out/r2/t_all.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2/bench1.json 2> gpurun_out/r2/bench1.err; echo bench $?; tail -5 gpurun_out/r2/bench1.err; cat gpurun_out/r2/bench1.json
