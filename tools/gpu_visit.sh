#!/bin/bash
# One GPU-box visit: parity tests, bench line, per-op breakdown.  usage: tools/gpu_visit.sh <tag> [pytest-args...]
O=gpurun_out/${1:-visit}; shift
mkdir -p $O
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $O/gpu.txt
timeout 1200 python -m pytest tests -m gpu -q "$@" > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log; tail -25 $O/pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench.json 2> $O/bench.err; tail -c 1700 $O/bench.json; tail -5 $O/bench.err
timeout 300 python tools/profile_step.py --batch 256 > $O/step_breakdown.txt 2>&1; head -64 $O/step_breakdown.txt
