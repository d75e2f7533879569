#!/bin/bash
# compute-sanitizer memcheck over the kernels added / changed in round 2 (small cases)
O=gpurun_out/r2san; mkdir -p $O
CS="compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20"
timeout 500 $CS python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "fused_assign or assignment_aggregation" > $O/memcheck_aggregate.log 2>&1; echo "rc=$?" >> $O/memcheck_aggregate.log; tail -4 $O/memcheck_aggregate.log
timeout 500 $CS python -m pytest tests/test_gemm_gpu.py -m gpu -q -x -k "tail" > $O/memcheck_gemm_tail.log 2>&1; echo "rc=$?" >> $O/memcheck_gemm_tail.log; tail -4 $O/memcheck_gemm_tail.log
timeout 600 $CS python -m pytest tests/test_e2e_gpu.py -m gpu -q -x -k "trained_stem and toy or mask_ratio and toy-0.5" > $O/memcheck_e2e.log 2>&1; echo "rc=$?" >> $O/memcheck_e2e.log; tail -4 $O/memcheck_e2e.log
grep -c "ERROR SUMMARY: 0 errors" $O/*.log
