#!/bin/bash
O=gpurun_out/r2m; mkdir -p $O
timeout 600 python -m pytest tests/test_attention_gpu.py tests/test_gemm_gpu.py -m gpu -q -x > $O/pytest_attn.log 2>&1; tail -6 $O/pytest_attn.log | cut -c1-300
timeout 120 python tools/one_attn.py vision; timeout 120 python tools/one_attn.py text
bash tools/gpu_visit.sh r2m
