mkdir -p gpurun_out/r2
timeout 300 python -m pytest tests/test_gpu_bm25_rrf.py tests/test_gpu_at_size.py -x -q -m gpu -k "bm25 or config3" > gpurun_out/r2/t_bm25.log 2>&1; echo tests $?; tail -3 gpurun_out/r2/t_bm25.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2/bench_1gpu.json 2> gpurun_out/r2/bench_1gpu.err; echo bench $?; python -c "
import json; d=json.load(open('gpurun_out/r2/bench_1gpu.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['stages_ms'], d['cpu_baseline']['gpu_matches_cpu_on_sample'], d['roofline']['frac'], d['roofline']['traffic'])"
