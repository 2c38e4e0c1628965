mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2/t_all.log 2>&1; echo tests $?; tail -6 gpurun_out/r2/t_all.log
