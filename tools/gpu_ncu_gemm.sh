#!/bin/bash
# ncu --set full captures (with SASS source view) of single GEMM shapes: $1 = tag, rest = bench_gemm name filters
O=gpurun_out/$1; shift
mkdir -p $O
i=0
for f in "$@"; do
  i=$((i+1))
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc2 -s 5 -c 1 -o $O/gemm_$i python tools/bench_gemm.py "$f" > $O/ncu_$i.log 2>&1
done
ls -la $O
