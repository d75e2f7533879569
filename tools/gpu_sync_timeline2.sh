#!/bin/bash
O=gpurun_out/r2sync; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29700"
timeout 300 $TR tools/sync_timeline.py > $O/sync_timeline_n2.txt 2> $O/err.txt; cat $O/sync_timeline_n2.txt | head -60; tail -5 $O/err.txt
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-240
