#!/bin/bash
# tail column slices of the 2-CTA GEMM: unit tests, micro-benchmark and step A/B on one box
O=gpurun_out/${1:-r2s}; mkdir -p $O
timeout 600 python -m pytest tests/test_gemm_gpu.py -m gpu -q -x > $O/pytest_gemm.log 2>&1; echo "rc=$?" >> $O/pytest_gemm.log; tail -6 $O/pytest_gemm.log
for t in 0 1; do
  echo "== SC_GEMM_TAIL_SPLIT=$t"; SC_GEMM_TAIL_SPLIT=$t timeout 300 python tools/bench_gemm.py 2>&1 | grep text | tee $O/gemm_microbench_tail$t.txt
done
timeout 900 python -m pytest tests/test_e2e_gpu.py -m gpu -q -x > $O/pytest_e2e.log 2>&1; echo "rc=$?" >> $O/pytest_e2e.log; tail -4 $O/pytest_e2e.log
for i in 1 2; do
  for p in 0 1; do
    SC_GEMM_TAIL_SPLIT=$p timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_tail${p}_$i.json 2> $O/bench_tail${p}_$i.err
    python - <<PY
import json
d=json.load(open("$O/bench_tail${p}_$i.json"))
print("TAIL=$p run $i: %.3f ms/step  %.0f pairs/s  e2e %.0f  fwd %.3f ms  gemm frac %.3f clocks %s" % (d["ms_per_step"], d["value"], d["e2e"]["value"], d["fwd_tensor_frac"]["ms_fwd"], d["roofline"]["frac"], d["clocks"]["sm_mhz"]))
PY
  done
done
