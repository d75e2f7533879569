#!/bin/bash
# quick GPU visit: $1 = output tag, $2 = pytest -k/paths (quoted), then GEMM micro-benchmark + bench line + breakdown
O=gpurun_out/${1:-quick}
mkdir -p $O
timeout 900 python -m pytest ${2:-tests} -m gpu -x -q 2>&1 | tail -15 > $O/pytest.log; tail -3 $O/pytest.log
timeout 300 python tools/bench_gemm.py > $O/gemm_microbench.txt 2>&1; cat $O/gemm_microbench.txt
timeout 300 python tools/one_attn.py > $O/attn.txt 2>&1; cat $O/attn.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench.json 2> $O/bench.err; python -c "
import json;d=json.load(open('$O/bench.json'));print({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')}, d['e2e']['value'], d['roofline']['achieved'], d['step_tensor_frac']['frac'])"
timeout 300 python tools/profile_step.py --batch 256 > $O/step_breakdown.txt 2>&1; head -12 $O/step_breakdown.txt
