#!/bin/bash
O=gpurun_out/r2pad; mkdir -p $O
timeout 300 python -m pytest tests/test_attention_gpu.py -m gpu -q -x > $O/pytest_attn.log 2>&1; echo "rc=$?" >> $O/pytest_attn.log; tail -6 $O/pytest_attn.log
timeout 600 python -m pytest tests/test_e2e_gpu.py -m gpu -q -x > $O/pytest_e2e.log 2>&1; echo "rc=$?" >> $O/pytest_e2e.log; tail -4 $O/pytest_e2e.log
for i in 1 2; do
  for p in 0 1; do
    SC_ATT_TC_PAD=$p timeout 600 python bench.py --heads --steps 8 --warmup 3 --no-cpu-baseline > $O/bench_heads_pad${p}_$i.json 2> $O/err.txt
    python - <<PY
import json
d=json.load(open("$O/bench_heads_pad${p}_$i.json"))
print("heads PAD=$p run $i: %.3f ms/step  %.0f pairs/s  clocks %s" % (d["ms_per_step"], d["value"], d["clocks"]["sm_mhz"]))
PY
  done
done
timeout 300 python tools/profile_step.py --heads --top 100 2>&1 | grep -E "^total|hd=48"
