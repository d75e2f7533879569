#!/bin/bash
# One GPU-box visit: parity tests, bench lines, per-op breakdown, ncu launch list, ncu --set full captures.
O=gpurun_out/${1:-r1b}
mkdir -p $O
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $O/gpu.txt
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -5 $O/pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > $O/bench.json 2> $O/bench.err; tail -c 1800 $O/bench.json
timeout 300 python bench.py --heads --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_heads.json 2>> $O/bench.err; tail -c 600 $O/bench_heads.json
timeout 300 python tools/profile_step.py --batch 256 > $O/step_breakdown.txt 2>&1
timeout 300 python tools/bench_gemm.py > $O/gemm_microbench.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/launches.csv python tools/profile_step.py --ncu > $O/ncu_launch.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file $O/launches_dram.csv python tools/profile_step.py --ncu > $O/ncu_launch2.log 2>&1
python tools/summarize_launches.py $O/launches.csv > $O/launch_summary.txt 2>&1
python tools/summarize_launches.py $O/launches_dram.csv --traffic $O/gemm_traffic.json > $O/launch_summary_dram.txt 2>&1
if [ "$2" != "nofull" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc2 -s 3 -c 1 -o $O/gemm_c_fc python tools/one_gemm.py c_fc > $O/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_tc -s 1 -c 1 -o $O/attn_bwd python tools/one_attn.py >> $O/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_tc -s 1 -c 1 -o $O/attn_fwd python tools/one_attn.py >> $O/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:ln_bwd -s 20 -c 1 -o $O/ln_bwd python tools/profile_step.py --ncu >> $O/ncu_full.log 2>&1
fi
ls -la $O
