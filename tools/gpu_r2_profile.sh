#!/bin/bash
# round 2 profiling visit: bench line, per-op breakdown, ncu launch lists (time; time + DRAM bytes) of one step, ncu --set full of
# the dominant kernels.  Numbers printed under ncu are never bench values.
O=gpurun_out/${1:-r2p}; mkdir -p $O
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $O/gpu.txt
timeout 600 python bench.py --steps 10 --warmup 3 > $O/bench.json 2> $O/bench.err; grep '^{' $O/bench.json | cut -c1-300
timeout 300 python tools/profile_step.py --batch 256 > $O/step_breakdown.txt 2>&1
python tools/roofline_by_kernel.py $O/step_breakdown.txt > $O/roofline_by_kernel.txt 2>&1 || true
timeout 300 python tools/bench_gemm.py > $O/gemm_microbench.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/launches.csv python tools/profile_step.py --ncu > $O/ncu_launch.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file $O/launches_dram.csv python tools/profile_step.py --ncu > $O/ncu_launch2.log 2>&1
python tools/summarize_launches.py $O/launches.csv > $O/launch_summary.txt 2>&1
python tools/summarize_launches.py $O/launches_dram.csv --traffic $O/gemm_traffic.json > $O/launch_summary_dram.txt 2>&1
cat $O/launch_summary.txt | head -20
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc2 -s 3 -c 1 -o $O/gemm_c_fc python tools/one_gemm.py c_fc > $O/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc2 -s 3 -c 1 -o $O/gemm_out_proj python tools/one_gemm.py out_proj >> $O/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_tc -s 1 -c 1 -o $O/attn_bwd python tools/one_attn.py >> $O/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_tc -s 1 -c 1 -o $O/attn_fwd python tools/one_attn.py >> $O/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:ln_bwd_tma -s 20 -c 1 -o $O/ln_bwd python tools/profile_step.py --ncu >> $O/ncu_full.log 2>&1
ls -la $O | head -30
