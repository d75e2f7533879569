#!/bin/bash
# final-build 2-GPU visit: multi-rank parity tests + one N=2 bench line
O=gpurun_out/r2m2_final; mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv > $O/gpu.txt
timeout 600 python -m pytest tests/test_multirank.py -m gpu -v > $O/pytest_multirank.log 2>&1; echo "pytest rc=$?" >> $O/pytest_multirank.log; tail -14 $O/pytest_multirank.log | cut -c1-220
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29700"
timeout 300 $TR bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err; grep '^{' $O/bench_n2.json | cut -c1-600
