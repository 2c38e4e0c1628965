#!/bin/bash
O=gpurun_out/r2/final; mkdir -p $O
timeout -k 10 400 ncu --set full --clock-control none --import-source on -k "regex:tc_tf32_scan_kernel" --launch-skip 4 -c 2 -f -o $O/exact_tc_scan python tools/exact_tc_profile.py > $O/ncu_exact2.log 2>&1; echo "ncu-exact $?"
tail -3 $O/ncu_exact2.log
