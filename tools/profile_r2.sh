#!/bin/bash
# Round-2 measurement pass on one B200 (run through gpurun): the bench line, the ncu launch list of
# the same command and one full capture of every kernel of the step.  Outputs under gpurun_out/r2/.
mkdir -p gpurun_out/r2
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2/bench_1gpu.json 2> gpurun_out/r2/bench_1gpu.err; echo bench $?; tail -3 gpurun_out/r2/bench_1gpu.err | cut -c1-300
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2/bench_ref.json 2> gpurun_out/r2/bench_ref.err; echo ref $?; cut -c1-400 gpurun_out/r2/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2/launches.csv python bench.py --profile --steps 2 --warmup 3 > gpurun_out/r2/launches.log 2>&1; echo launches $?
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'bm25_fast_kernel|tc_i8_search_kernel|rescore_ring_kernel|tau_keys_kernel|bm25_refine_kernel|tc_select_lists_kernel|rank_scored|rrf_fuse|quantize_ubinary|unpack_pm1' --launch-skip 30 -c 15 -o gpurun_out/r2/step_full -f python bench.py --profile --steps 1 --warmup 3 > gpurun_out/r2/ncu_full.log 2>&1; echo full $?; tail -2 gpurun_out/r2/ncu_full.log
ls -la gpurun_out/r2 | tail -8
