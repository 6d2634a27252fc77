timeout 300 python -m pytest tests -m gpu -x -q -k "tensor_core" 2>&1 | tail -3
for i in 1 2; do
timeout 200 python bench.py --steps 100 --warmup 5 --no-cpu 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('run ms', round(d['roofline']['kernel_ms'],4), d['roofline']['frac'], d['value'], d['e2e']['value'])"
done
