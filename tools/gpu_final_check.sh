#!/bin/bash
# what the driver runs at round end: smoke, pytest -m gpu, bench (own arm with the CPU baseline), bench --impl reference
O=gpurun_out/r2check; mkdir -p $O
timeout 600 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log | cut -c1-400
timeout 1200 python -m pytest tests -x -q -m gpu > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest.log
timeout 600 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; cut -c1-500 $O/bench.json; grep -o '"cpu_baseline".*' $O/bench.json | cut -c1-300
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$?"; cut -c1-300 $O/bench_ref.json
