mkdir -p gpurun_out/r2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --debug-extra config5 --debug-rows 8000000 > gpurun_out/r2/cfg5_dbg.json 2> gpurun_out/r2/cfg5_dbg.err; echo cfg5 $?; tail -5 gpurun_out/r2/cfg5_dbg.err | cut -c1-400; cat gpurun_out/r2/cfg5_dbg.json | cut -c1-3000
