#!/bin/bash
# Final single-GPU validation of a round: tests, smoke, bench (both arms), peaks, probes, ncu captures.
O=gpurun_out/r2/final; mkdir -p $O
timeout -k 10 1200 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest $?"; tail -3 $O/pytest_gpu.log | cut -c1-300
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; echo "smoke $?"; tail -2 $O/smoke.log | cut -c1-300
timeout -k 10 900 python bench.py --steps 20 --warmup 5 > $O/bench_1.json 2> $O/bench_1.err; echo "bench $?"; tail -2 $O/bench_1.err | cut -c1-300
timeout -k 10 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref $?"
timeout -k 10 200 python tools/peak_probe.py $O/peaks.json > /dev/null 2> $O/peaks.err; echo "peaks $?"
timeout -k 10 300 python tools/exact_probe.py --f32-only > $O/exact_probe.jsonl 2> $O/exact_probe.err; echo "exact_probe $?"
K="regex:rr::|tc_i8|bm25_|tau_keys|rescore_ring|tc_select|rank_scored|rrf_fuse|quantize_|unpack_pm1"
timeout -k 10 400 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" --launch-skip 35 -c 13 --csv --log-file $O/launches_step.csv python tools/step_profile.py 3 > $O/ncu_launch.log 2>&1; echo "ncu-launch $?"
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k "$K" --launch-skip 35 -c 13 -f -o $O/step_full python tools/step_profile.py 3 > $O/ncu_full.log 2>&1; echo "ncu-full $?"
timeout -k 10 400 ncu --set full --clock-control none --import-source on -k "regex:tc_tf32|tx_refine|tc_select_lists|tx_qnorm|tau_keys" --launch-skip 10 -c 5 -f -o $O/exact_tc_full python tools/exact_tc_profile.py > $O/ncu_exact.log 2>&1; echo "ncu-exact $?"
timeout -k 10 400 ncu --set full --clock-control none --import-source on -k "regex:tc_tf32_scan_kernel" --launch-skip 4 -c 2 -f -o $O/exact_tc_scan python tools/exact_tc_profile.py > $O/ncu_exact2.log 2>&1; echo "ncu-exact-scan $?"
python -c "
import json; d=json.load(open('$O/bench_1.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['cpu_baseline']['gpu_matches_cpu_on_sample'], d['roofline']['frac']); print({k:(v.get('value'), v.get('parity_on_sample') or (v.get('cpu_baseline') or {}).get('gpu_matches_cpu_on_sample')) for k,v in d['extras'].items()})
print(open('$O/bench_ref.json').read()[:400])"
ls -la $O
