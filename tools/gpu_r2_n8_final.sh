#!/bin/bash
# final-build 8-GPU visit: W=8 multi-rank parity, BASELINE configs 2/3/4/5 at N=8, gradient-sync timeline
O=gpurun_out/r2n8_final; mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv > $O/gpu.txt
timeout 400 python -m pytest tests/test_multirank.py -m gpu -v -k "8-native or allreduce_matches_nccl and 8" > $O/pytest_multirank_w8.log 2>&1; echo "pytest rc=$?" >> $O/pytest_multirank_w8.log; tail -8 $O/pytest_multirank_w8.log | cut -c1-200
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29700"
run() { name=$1; shift; timeout 300 $TR bench.py --gpus 8 --steps 8 --warmup 3 "$@" > $O/$name.json 2> $O/$name.err; grep '^{' $O/$name.json | cut -c1-330; }
run bench_cfg2_b256_n8_nvls
run bench_cfg3_b512_n8 --batch 512
run bench_cfg4_heads_n8 --heads
run bench_cfg5_vitl14_n8 --model vitl14
timeout 200 $TR tools/sync_timeline.py > $O/sync_timeline_n8.txt 2> $O/sync_timeline.err; grep -E "EXPOSED|exposed|one step" $O/sync_timeline_n8.txt
