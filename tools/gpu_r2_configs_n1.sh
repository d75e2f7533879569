#!/bin/bash
# BASELINE configs 3 / 4 / 5 at N = 1 (per-GPU shapes) + the config-5 block sweep; lines -> gpurun_out/$1
O=gpurun_out/${1:-r2cfg}; mkdir -p $O
timeout 600 python bench.py --batch 512 --steps 6 --warmup 3 --no-cpu-baseline > $O/bench_cfg3_b512_n1.json 2> $O/err.txt; tail -c 900 $O/bench_cfg3_b512_n1.json
timeout 600 python bench.py --heads --steps 6 --warmup 3 --no-cpu-baseline > $O/bench_cfg4_heads_n1.json 2>> $O/err.txt; tail -c 900 $O/bench_cfg4_heads_n1.json
timeout 600 python bench.py --model vitl14 --steps 6 --warmup 3 --no-cpu-baseline > $O/bench_cfg5_vitl14_n1.json 2>> $O/err.txt; tail -c 900 $O/bench_cfg5_vitl14_n1.json
timeout 600 python tools/block_sweep.py --batch 64 > $O/block_sweep_vitl.txt 2>> $O/err.txt; cat $O/block_sweep_vitl.txt
timeout 600 python tools/block_sweep.py --batch 256 --width 768 --seq 196 > $O/block_sweep_vitb.txt 2>> $O/err.txt; cat $O/block_sweep_vitb.txt
timeout 300 python tools/profile_step.py --heads --top 30 > $O/step_breakdown_heads.txt 2>> $O/err.txt; head -45 $O/step_breakdown_heads.txt
timeout 300 python tools/profile_step.py --model vitl14 --top 30 > $O/step_breakdown_vitl14.txt 2>> $O/err.txt; head -45 $O/step_breakdown_vitl14.txt
tail -20 $O/err.txt
