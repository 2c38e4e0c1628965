mkdir -p gpurun_out/r2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2/bench2.json 2> gpurun_out/r2/bench2.err; echo bench2 $?; tail -5 gpurun_out/r2/bench2.err | cut -c1-300; cat gpurun_out/r2/bench2.json | cut -c1-6000
