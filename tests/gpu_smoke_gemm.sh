#!/bin/bash
# first-contact run for the tcgen05 GEMM: bounded by timeout so a protocol bug cannot hang the box
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/gpu.txt
timeout 600 python -m pytest tests/test_gemm_gpu.py -x -q -m gpu 2>&1 | tail -40 | tee gpurun_out/gemm_test.log
