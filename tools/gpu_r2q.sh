#!/bin/bash
# A/B of programmatic dependent launch on one box: tests with PDL on, bench with SC_PDL=0 / 1 (twice, interleaved).
O=gpurun_out/${1:-r2q}; mkdir -p $O
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $O/gpu.txt
timeout 1200 python -m pytest tests -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log; tail -8 $O/pytest.log
for i in 1 2; do
  for p in 0 1; do
    SC_PDL=$p timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_pdl${p}_$i.json 2> $O/bench_pdl${p}_$i.err
    python - <<PY
import json
d=json.load(open("$O/bench_pdl${p}_$i.json"))
print("PDL=$p run $i: %.3f ms/step  %.0f pairs/s  e2e %.0f  fwd %.3f ms  clocks %s" % (d["ms_per_step"], d["value"], d["e2e"]["value"], d["fwd_tensor_frac"]["ms_fwd"], d["clocks"]))
PY
    tail -3 $O/bench_pdl${p}_$i.err
  done
done
