#!/bin/bash
# round 2, GPU visit 1: parity suite with the production-dispatch cases, attention timelines, TMEM / MUFU micro-benchmark,
# baseline bench line + per-op breakdown of the unchanged kernels
O=gpurun_out/r2a
mkdir -p $O
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $O/gpu.txt; nproc >> $O/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log; tail -5 $O/pytest.log
timeout 120 tools/ubench/tmem_bw > $O/tmem_bw.txt 2>&1; cat $O/tmem_bw.txt
for w in vision text; do for p in bwd fwd; do
  SEGCLIP_B200_LIB=segclip_b200/lib_trace/libsegclip_b200.so timeout 120 python tools/trace_attn.py $w $p > $O/trace_${w}_${p}.txt 2>&1
done; done
head -60 $O/trace_vision_bwd.txt
timeout 300 python tools/e2e_report.py vitb16:8 bf16 forced > $O/parity_b8_contrastive.txt 2>&1; head -3 $O/parity_b8_contrastive.txt | cut -c1-900
timeout 300 python tools/e2e_report.py vitb16:16:heads bf16 forced > $O/parity_b16_heads.txt 2>&1; head -3 $O/parity_b16_heads.txt | cut -c1-900
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench.json 2> $O/bench.err; tail -c 1500 $O/bench.json
timeout 300 python tools/profile_step.py --batch 256 > $O/step_breakdown.txt 2>&1; head -30 $O/step_breakdown.txt
