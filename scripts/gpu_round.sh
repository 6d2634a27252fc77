#!/bin/bash
# One GPU session: parity tests, smoke, bench (ours + reference arm), ncu launch list + full capture.
set -x
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 300 python bench.py --steps 50 --warmup 5 2>&1 | tail -1 > gpurun_out/bench_n1.json; cat gpurun_out/bench_n1.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_ref.json; cat gpurun_out/bench_ref.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_c2.csv python bench.py --steps 5 --warmup 3 --no-cpu > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_fused -s 3 -c 1 -o gpurun_out/prof_fused_c2 python bench.py --steps 2 --warmup 3 --no-cpu > /dev/null 2>&1
ls -la gpurun_out | tail -5
