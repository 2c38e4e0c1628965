#!/bin/bash
mkdir -p gpurun_out/r2
N=$1
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 --no-extras > gpurun_out/r2/bench_ov_$N.json 2> gpurun_out/r2/bench_ov_$N.err
echo "overlap $?"; tail -3 gpurun_out/r2/bench_ov_$N.err | cut -c1-300; python -c "
import json; d=json.load(open('gpurun_out/r2/bench_ov_$N.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['exchange_us'], d['cpu_baseline']['gpu_matches_cpu_on_sample'])"
timeout -k 10 300 python -m pytest tests/test_gpu_at_size.py -x -q -k "config3" 2>&1 | tail -5
