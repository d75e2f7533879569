#!/bin/bash
O=gpurun_out/${1:-r2x}; mkdir -p $O
timeout 900 python -m pytest tests/test_attention_gpu.py tests/test_e2e_gpu.py -m gpu -q -x > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log; tail -5 $O/pytest.log
timeout 300 python tools/profile_step.py --batch 256 --top 80 > $O/step_breakdown.txt 2>&1; grep -E "^total|Lq=8|attention" $O/step_breakdown.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench.json 2> $O/bench.err
python - <<PY
import json
d=json.load(open("$O/bench.json"))
print("%.3f ms/step  %.0f pairs/s  e2e %.0f  fwd %.3f ms  clocks %s" % (d["ms_per_step"], d["value"], d["e2e"]["value"], d["fwd_tensor_frac"]["ms_fwd"], d["clocks"]["sm_mhz"]))
PY
