#!/bin/bash
# usage: gpu_bench_n.sh N   -> gpurun_out/r2/bench_final_N.json
N=$1
mkdir -p gpurun_out/r2
if [ "$N" = "1" ]; then
  timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2/bench_final_1.json 2> gpurun_out/r2/bench_final_1.err
else
  timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2/bench_final_$N.json 2> gpurun_out/r2/bench_final_$N.err
fi
echo bench$N $?; tail -3 gpurun_out/r2/bench_final_$N.err | cut -c1-300
python -c "
import json; d=json.load(open('gpurun_out/r2/bench_final_$N.json')); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['exchange_us'], d['cpu_baseline']['gpu_matches_cpu_on_sample']); print({k:(v.get('value'), v.get('parity_on_sample') or (v.get('cpu_baseline') or {}).get('gpu_matches_cpu_on_sample')) for k,v in d['extras'].items()})"
