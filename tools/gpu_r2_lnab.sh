#!/bin/bash
O=gpurun_out/r2lnab; mkdir -p $O
for i in 1 2; do
  for d in 640 512; do
    SC_LN_TMA_MIN_D=$d timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_d${d}_$i.json 2> $O/err.txt
    python - <<PY
import json
d=json.load(open("$O/bench_d${d}_$i.json"))
print("LN_TMA_MIN_D=$d run $i: %.3f ms/step  %.0f pairs/s fwd %.3f clocks %s" % (d["ms_per_step"], d["value"], d["fwd_tensor_frac"]["ms_fwd"], d["clocks"]["sm_mhz"]))
PY
  done
done
SC_LN_TMA_MIN_D=512 timeout 300 python tools/profile_step.py --batch 256 --top 100 2>&1 | grep -E "^total|layernorm_bwd"
timeout 300 python tools/profile_step.py --batch 256 --top 100 2>&1 | grep -E "^total|layernorm_bwd"
