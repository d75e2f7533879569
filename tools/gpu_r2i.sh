#!/bin/bash
O=gpurun_out/r2j; mkdir -p $O
timeout 600 python -m pytest tests/test_attention_gpu.py -m gpu -q -x > $O/pytest_attn.log 2>&1; tail -6 $O/pytest_attn.log | cut -c1-300
timeout 120 python tools/one_attn.py > $O/one_attn.txt 2>&1; cat $O/one_attn.txt
for w in vision text; do SEGCLIP_B200_LIB=segclip_b200/lib_trace/libsegclip_b200.so timeout 120 python tools/trace_attn.py $w bwd > $O/trace_${w}_bwd.txt 2>&1; done
grep -n "softmax warp" -A40 $O/trace_vision_bwd.txt | head -50
bash tools/gpu_visit.sh r2j
