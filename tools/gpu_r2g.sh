#!/bin/bash
O=gpurun_out/r2g; mkdir -p $O
timeout 600 python -m pytest tests/test_attention_gpu.py -m gpu -q -x > $O/pytest_attn.log 2>&1; tail -6 $O/pytest_attn.log | cut -c1-300
timeout 120 python tools/one_attn.py > $O/one_attn.txt 2>&1; cat $O/one_attn.txt
SC_ATT_FWD_V1=1 timeout 120 python tools/one_attn.py > $O/one_attn_v1.txt 2>&1; cat $O/one_attn_v1.txt
bash tools/gpu_visit.sh r2g
