#!/bin/bash
# usage: gpu_bench_n.sh N [extra bench flags]  -> gpurun_out/r2/final/bench_N.json
N=$1; shift
O=gpurun_out/r2/final; mkdir -p $O
if [ "$N" = "1" ]; then
  timeout -k 10 900 python bench.py --gpus 1 --steps 20 --warmup 5 "$@" > $O/bench_1.json 2> $O/bench_1.err
else
  timeout -k 10 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 5 "$@" > $O/bench_$N.json 2> $O/bench_$N.err
fi
echo bench$N $?; tail -3 $O/bench_$N.err | cut -c1-300
python -c "
import json; d=json.load(open('$O/bench_$N.json')); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['exchange_us'], d['cpu_baseline']['gpu_matches_cpu_on_sample']); print({k:(v.get('value'), v.get('ms_per_step'), v.get('parity_on_sample') or (v.get('cpu_baseline') or {}).get('gpu_matches_cpu_on_sample')) for k,v in d.get('extras', {}).items()})"
