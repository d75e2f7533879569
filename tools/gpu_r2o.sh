#!/bin/bash
O=gpurun_out/r2o; mkdir -p $O
timeout 600 python -m pytest tests/test_gemm_gpu.py -m gpu -q -x > $O/pytest_gemm.log 2>&1; tail -6 $O/pytest_gemm.log | cut -c1-300
timeout 300 python tools/bench_gemm.py > $O/gemm_microbench.txt 2>&1; grep "c_fc \|dgrad\*\|text c_fc\|text c_proj dgrad" $O/gemm_microbench.txt
bash tools/gpu_visit.sh r2o
