#!/bin/bash
run() { echo "== $1 heads=$2"; RR_B200_LIB=$PWD/radiant-rag_b200/librr_b200$1.so timeout -k 10 200 python tools/bm25_probe.py 1000000 1024 5 1024 $2 2>&1 | tail -1 | cut -c150-400; }
run _if4 64
run _if6 64
run _if8 64
run _if10 64
run _if6 64
