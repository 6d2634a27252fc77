#!/bin/bash
# One GPU session: parity tests, smoke, bench lines (ours + reference arm, C3, Gibbs, C4 shard), ncu launch lists
# and full captures of the three hot kernels.  Outputs land in gpurun_out/; scripts/summarize_ncu.py turns them
# into the tracked summaries under profiles/.
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 300 python bench.py --steps 50 --warmup 5 2>&1 | tail -1 > gpurun_out/bench_n1.json; cat gpurun_out/bench_n1.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_ref.json; cat gpurun_out/bench_ref.json
timeout 300 python bench.py --workload c3 --steps 10 --warmup 3 --no-cpu 2>&1 | tail -1 > gpurun_out/bench_c3.json; cut -c1-200 gpurun_out/bench_c3.json
timeout 300 python bench.py --workload c3-gibbs --steps 10 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_c3_gibbs.json; cut -c1-200 gpurun_out/bench_c3_gibbs.json
timeout 300 python bench.py --workload c4-shard --steps 5 --warmup 3 --no-cpu 2>&1 | tail -1 > gpurun_out/bench_c4_shard.json; cut -c1-200 gpurun_out/bench_c4_shard.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_c2.csv python bench.py --steps 5 --warmup 3 --no-cpu > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_fused -s 3 -c 1 -f -o gpurun_out/prof_fused_c2 python bench.py --steps 2 --warmup 3 --no-cpu > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gibbs_delta -s 2 -c 1 -f -o gpurun_out/prof_gibbs_c3 python bench.py --workload c3-gibbs --steps 2 --warmup 3 --no-cpu > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none -k regex:tc_gemm -c 12 --csv --log-file gpurun_out/launches_c3_gemm.csv python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu > /dev/null 2>&1
ls -la gpurun_out | tail -12
