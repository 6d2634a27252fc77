#!/bin/bash
# compute-sanitizer passes over the small GPU tests (run under gpurun).  memcheck: every kernel, including the
# TMA / tcgen05 ones; racecheck: the kernels that synchronise with bar.sync only (K1, FP64 path, K4).
set -x
timeout 1600 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 3 \
    python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_inference.py -m gpu -q -k "not benchmark_size" 2>&1 | grep -v "Host Frame" | tail -8
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 5 \
    python -m pytest tests/test_gpu_parity.py -m gpu -q -k "filter or fp64_path or gibbs or dense or firing" 2>&1 | grep -v "Host Frame" | tail -6
