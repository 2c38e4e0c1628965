mkdir -p gpurun_out/r2
P=radiant-rag_b200
for cfg in "librr_b200.so" "librr_b200_st16.so" "librr_b200_st20.so"; do
  RR_B200_LIB=$PWD/$P/$cfg timeout 200 python tools/bm25_probe.py 1000000 1024 5 2>&1 | tail -1
done
RR_B200_LIB=$PWD/$P/librr_b200_st16.so timeout 300 python -m pytest tests/test_gpu_bm25_rrf.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python -m pytest tests/test_gpu_bm25_rrf.py tests/test_gpu_at_size.py -x -q -m gpu -k "bm25 or config3" 2>&1 | tail -3
