#!/bin/bash
mkdir -p gpurun_out/r2
timeout -k 10 500 python -m pytest tests/test_gpu_rescore.py tests/test_gpu_flow.py -x -q 2>&1 | tail -3
timeout -k 10 200 python tools/peak_probe.py gpurun_out/r2/peaks.json | cut -c1-1200
timeout -k 10 300 python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/r2/bench_rs.json 2> gpurun_out/r2/bench_rs.err
echo "bench $?"; tail -3 gpurun_out/r2/bench_rs.err | cut -c1-300; python -c "
import json; d=json.load(open('gpurun_out/r2/bench_rs.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['cpu_baseline']['gpu_matches_cpu_on_sample']); print(json.dumps(d.get('stages_ms', d.get('stage_ms', '')))[:600])"
