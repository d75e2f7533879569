#!/bin/bash
# round 2, 2-GPU visit: multi-rank parity (P2P exchange, NCCL comparator, native gradient sync, rank-0-only eval) + N=2 A/B of
# the NVLink P2P exchange against the NCCL all-gather comparator
O=gpurun_out/r2m2
mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv > $O/gpu.txt
timeout 600 python -m pytest tests/test_multirank.py -m gpu -v > $O/pytest_multirank.log 2>&1; echo "pytest rc=$?" >> $O/pytest_multirank.log; tail -12 $O/pytest_multirank.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29700"
timeout 300 $TR bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_n2_p2p.json 2> $O/bench_n2_p2p.err; tail -c 1200 $O/bench_n2_p2p.json
SEGCLIP_EXCHANGE=nccl timeout 300 $TR bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_n2_nccl.json 2> $O/bench_n2_nccl.err; tail -c 1200 $O/bench_n2_nccl.json
