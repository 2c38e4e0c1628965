#!/bin/bash
timeout -k 10 300 python -m pytest tests/test_gpu_bm25_rrf.py -x -q 2>&1 | tail -3
for v in "" _rf; do
  echo "variant '$v'"; RR_B200_LIB=$PWD/radiant-rag_b200/librr_b200$v.so timeout -k 10 200 python tools/bm25_probe.py 1000000 1024 5 2>&1 | tail -1 | cut -c120-420
done
RR_B200_LIB=$PWD/radiant-rag_b200/librr_b200.so timeout -k 10 200 python tools/bm25_probe.py 125000 1024 5 2>&1 | tail -1 | cut -c120-420
