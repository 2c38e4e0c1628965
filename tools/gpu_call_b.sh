#!/bin/bash
mkdir -p gpurun_out/r2
timeout -k 10 500 python -m pytest tests/test_gpu_bm25_rrf.py tests/test_gpu_at_size.py -x -q 2>&1 | tail -5 | cut -c1-300
timeout -k 10 300 python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/r2/bench_w32.json 2> gpurun_out/r2/bench_w32.err
echo "bench $?"; tail -3 gpurun_out/r2/bench_w32.err | cut -c1-300; python -c "
import json; d=json.load(open('gpurun_out/r2/bench_w32.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['cpu_baseline']['gpu_matches_cpu_on_sample']); print(json.dumps(d.get('stages_ms', d.get('stage_ms', '')))[:600]); print(d['roofline']['frac'], d['roofline'].get('smem_view'))"
