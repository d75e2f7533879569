#!/bin/bash
# 2-GPU visit: NVLS gradient-sync kernel (unit + end-to-end parity) and N=2 A/B of NVLS vs NCCL bucket all-reduce
O=gpurun_out/r2m2b; mkdir -p $O
timeout 900 python -m pytest tests/test_multirank.py -m gpu -v -x > $O/pytest_multirank.log 2>&1; echo "pytest rc=$?" >> $O/pytest_multirank.log; tail -25 $O/pytest_multirank.log | cut -c1-300
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29700"
timeout 300 $TR bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_n2_nvls.json 2> $O/bench_n2_nvls.err; tail -c 700 $O/bench_n2_nvls.json; tail -5 $O/bench_n2_nvls.err
SEGCLIP_GRAD_SYNC=nccl timeout 300 $TR bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_n2_nccl.json 2> $O/bench_n2_nccl.err; tail -c 700 $O/bench_n2_nccl.json
