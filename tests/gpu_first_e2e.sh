#!/bin/bash
mkdir -p gpurun_out
for c in toy_contrastive_flat toy_heads_flat toy_heads_per_sample; do
  echo "=== $c fp32" ; timeout 300 python tools/e2e_report.py $c fp32 2>&1 | tail -25
done 2>&1 | tee gpurun_out/e2e_first.log
