#!/bin/bash
O=gpurun_out/r2f; mkdir -p $O
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k layernorm > $O/pytest_ln.log 2>&1; tail -4 $O/pytest_ln.log
timeout 300 python tools/bench_ln.py > $O/ln_microbench.txt 2>&1; cat $O/ln_microbench.txt
SC_LN_BWD_REGS=1 timeout 300 python tools/bench_ln.py > $O/ln_microbench_regs.txt 2>&1; grep "ln_bwd bf16" $O/ln_microbench_regs.txt
bash tools/gpu_visit.sh r2f
