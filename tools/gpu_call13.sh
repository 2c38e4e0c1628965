mkdir -p gpurun_out/r2
timeout 120 python tools/step_profile.py 3 > gpurun_out/r2/step_profile.log 2>&1; tail -1 gpurun_out/r2/step_profile.log
# matching launches before the third step: 8 quantize (dense build) + 1 impacts + 2 x 13 step kernels
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'quantize_ubinary|unpack_pm1|tc_i8_search|tau_keys|tc_select_lists|rescore_ring|rank_scored|bm25_fast|bm25_refine|rrf_fuse|bm25_impacts' --launch-skip 35 -c 13 -o gpurun_out/r2/step_full2 -f python tools/step_profile.py 3 > gpurun_out/r2/ncu_full2.log 2>&1; echo full $?; grep -c Profiling gpurun_out/r2/ncu_full2.log; tail -2 gpurun_out/r2/ncu_full2.log
