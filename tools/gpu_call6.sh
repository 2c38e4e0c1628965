mkdir -p gpurun_out/r2
P=radiant-rag_b200
for cfg in "librr_b200.so 1024" "librr_b200.so 512" "librr_b200_w32.so 1024" "librr_b200_w32.so 512" "librr_b200_w16.so 1024"; do
  set -- $cfg
  RR_B200_LIB=$PWD/$P/$1 timeout 200 python tools/bm25_probe.py 1000000 1024 5 $2 2>&1 | tail -1
done
timeout 300 python -m pytest tests/test_gpu_bm25_rrf.py -x -q -m gpu 2>&1 | tail -3
