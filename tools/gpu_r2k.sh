#!/bin/bash
O=gpurun_out/r2k; mkdir -p $O
timeout 600 python -m pytest tests/test_seg_inference.py tests/test_e2e_gpu.py -m gpu -q -x -k "seg or bicubic or eval or inference" > $O/pytest_seg.log 2>&1; tail -8 $O/pytest_seg.log | cut -c1-300
timeout 300 python tools/profile_step.py --batch 256 --top 60 > $O/step_breakdown.txt 2>&1; grep "layernorm_bwd\|^total" $O/step_breakdown.txt
SC_LN_TMA_MIN_D=512 timeout 300 python tools/profile_step.py --batch 256 --top 60 > $O/step_breakdown_lntma512.txt 2>&1; grep "layernorm_bwd\|^total" $O/step_breakdown_lntma512.txt
