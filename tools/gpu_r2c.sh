#!/bin/bash
O=gpurun_out/r2c; mkdir -p $O
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_ops_gpu.py -m gpu -q -x > $O/pytest_units.log 2>&1; tail -5 $O/pytest_units.log
timeout 300 python tools/bench_gemm.py > $O/gemm_microbench.txt 2>&1; cat $O/gemm_microbench.txt
timeout 300 python tools/bench_ln.py > $O/ln_microbench.txt 2>&1; cat $O/ln_microbench.txt
SC_LN_BWD_SMEM=1 timeout 300 python tools/bench_ln.py > $O/ln_microbench_old.txt 2>&1; cat $O/ln_microbench_old.txt
bash tools/gpu_visit.sh r2c
