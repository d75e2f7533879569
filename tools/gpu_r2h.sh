#!/bin/bash
O=gpurun_out/r2h; mkdir -p $O
timeout 600 python -m pytest tests/test_attention_gpu.py -m gpu -q -x > $O/pytest_attn.log 2>&1; tail -6 $O/pytest_attn.log | cut -c1-300
timeout 120 python tools/one_attn.py > $O/one_attn.txt 2>&1; cat $O/one_attn.txt
SC_ATT_FWD_TWO_PASS=1 timeout 120 python tools/one_attn.py > $O/one_attn_twopass.txt 2>&1; cat $O/one_attn_twopass.txt
SEGCLIP_B200_LIB=segclip_b200/lib_trace/libsegclip_b200.so timeout 120 python tools/trace_attn.py vision fwd > $O/trace_vision_fwd.txt 2>&1
grep -n "softmax warp" -A16 $O/trace_vision_fwd.txt
bash tools/gpu_visit.sh r2h
