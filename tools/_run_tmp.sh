set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_modules.py -m gpu -x -q -s -k "skip_propagation" 2>&1 | tail -12
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-train 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value']); print(d['skip_propagation'])"
