PYGLM_NVCC_EXTRA="-DPYGLM_TC_ROLE_BASE=1" python theano_pyglm_b200/build.py > /dev/null
export PYGLM_NVCC_EXTRA="-DPYGLM_TC_ROLE_BASE=1"
timeout 300 python -m pytest tests -m gpu -x -q -k "tensor_core" 2>&1 | tail -3
timeout 200 python bench.py --steps 100 --warmup 5 --no-cpu 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('role-base-1 ms', round(d['roofline']['kernel_ms'],4), d['roofline']['frac'], d['value'], d['e2e']['value'])"
PYGLM_TC_TRACE=1 timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu 2>&1 | grep -v "^{" | grep "r_ready arrivals"
