set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_modules.py -m gpu -x -q -s 2>&1 | tail -40 > gpurun_out/r2f_tests.log
tail -15 gpurun_out/r2f_tests.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
cat gpurun_out/r2f_bench.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(json.dumps(d['train'],indent=1)); print(d['value'], d['e2e'])"
tail -5 gpurun_out/r2f_bench.err
