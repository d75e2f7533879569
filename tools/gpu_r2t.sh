#!/bin/bash
# fused aggregation kernels: unit + e2e tests, per-op times, step A/B on one box
O=gpurun_out/${1:-r2t}; mkdir -p $O
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_e2e_gpu.py tests/test_seg_inference.py -m gpu -q -x > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log; tail -12 $O/pytest.log
for t in 1 0; do
  if [ $t = 1 ]; then export SC_AGG_UNFUSED=1; else unset SC_AGG_UNFUSED; fi
  timeout 300 python tools/profile_step.py --batch 256 > $O/step_breakdown_unfused$t.txt 2>&1; grep -E "^total|assign|aggregate" $O/step_breakdown_unfused$t.txt | head -8
done
for i in 1 2; do
  for t in 1 0; do
    if [ $t = 1 ]; then export SC_AGG_UNFUSED=1; else unset SC_AGG_UNFUSED; fi
    timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_unfused${t}_$i.json 2> $O/bench_unfused${t}_$i.err
    python - <<PY
import json
d=json.load(open("$O/bench_unfused${t}_$i.json"))
print("UNFUSED=$t run $i: %.3f ms/step  %.0f pairs/s  e2e %.0f  fwd %.3f ms  clocks %s" % (d["ms_per_step"], d["value"], d["e2e"]["value"], d["fwd_tensor_frac"]["ms_fwd"], d["clocks"]["sm_mhz"]))
PY
  done
done
