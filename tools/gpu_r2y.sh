#!/bin/bash
O=gpurun_out/${1:-r2y}; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q -x > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log; tail -30 $O/pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench.json 2> $O/bench.err
python - <<PY
import json
d=json.load(open("$O/bench.json"))
print("%.3f ms/step  %.0f pairs/s  e2e %.0f  fwd %.3f ms  clocks %s" % (d["ms_per_step"], d["value"], d["e2e"]["value"], d["fwd_tensor_frac"]["ms_fwd"], d["clocks"]["sm_mhz"]))
PY
timeout 300 python tools/profile_step.py --batch 256 --top 100 > $O/step_breakdown.txt 2>&1; grep -E "^total|simt" $O/step_breakdown.txt
