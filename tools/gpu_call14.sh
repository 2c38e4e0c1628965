mkdir -p gpurun_out/r2
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2/bench8.json 2> gpurun_out/r2/bench8.err; echo bench8 $?; tail -5 gpurun_out/r2/bench8.err | cut -c1-400; cut -c1-1500 gpurun_out/r2/bench8.json
