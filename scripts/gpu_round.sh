#!/bin/bash
# One GPU session (gpurun, one B200): parity tests, smoke, bench lines (ours + reference arm), ncu launch lists and full
# captures of the hot kernels.  Outputs land in gpurun_out/; scripts/summarize_ncu.py turns them into the tracked
# summaries under profiles/ (see profiles/README.md for the command behind every file).
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 600 python bench.py 2>/dev/null | tail -1 > gpurun_out/bench_n1.json; cut -c1-300 gpurun_out/bench_n1.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_ref.json; cut -c1-300 gpurun_out/bench_ref.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_launches_c2.csv python bench.py --steps 5 --warmup 3 --no-cpu --no-extra > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_fused -s 3 -c 1 -f -o gpurun_out/r2_prof_fused_c2 python bench.py --steps 2 --warmup 3 --no-cpu --no-extra > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:filter_kernel -s 2 -c 1 -f -o gpurun_out/r2_prof_filter_planes python scripts/filter_bench.py planes > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:filter_kernel -s 2 -c 1 -f -o gpurun_out/r2_prof_filter_f32 python scripts/filter_bench.py f32 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gibbs_delta -s 2 -c 1 -f -o gpurun_out/r2_prof_gibbs_c3 python bench.py --workload c3-gibbs --steps 2 --warmup 3 --no-cpu > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none -k regex:tc_gemm -c 12 --csv --log-file gpurun_out/r2_launches_c3_gemm.csv python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu --no-extra > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_launches_fromspikes.csv python bench.py --workload c4-neuron-shard --steps 2 --warmup 1 --no-cpu --no-extra > /dev/null 2>&1
ls -la gpurun_out | tail -12
