mkdir -p gpurun_out/r2
timeout 400 python -m pytest tests/test_gpu_bm25_rrf.py tests/test_gpu_rescore.py -x -q -m gpu > gpurun_out/r2/t_bm25.log 2>&1; echo tests $?; tail -5 gpurun_out/r2/t_bm25.log
timeout 200 python tools/bm25_probe.py > gpurun_out/r2/bm25_probe.log 2>&1; tail -1 gpurun_out/r2/bm25_probe.log
timeout 200 python tools/bm25_probe.py 125000 1024 5 > gpurun_out/r2/bm25_probe125.log 2>&1; tail -1 gpurun_out/r2/bm25_probe125.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-extras > gpurun_out/r2/bench1b.json 2> gpurun_out/r2/bench1b.err; echo bench $?; python -c "
import json; d=json.load(open('gpurun_out/r2/bench1b.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['stages_ms'], d['cpu_baseline']['gpu_matches_cpu_on_sample'])"
