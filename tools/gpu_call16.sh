mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests -x -q -m gpu --durations=5 > gpurun_out/r2/t_all.log 2>&1; echo tests $?; tail -14 gpurun_out/r2/t_all.log
timeout 100 python __graft_entry__.py smoke 2>&1 | tail -1
