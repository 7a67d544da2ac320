set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mesh.py -m gpu -x -q -s 2>&1 | tail -40 > gpurun_out/r2j_tests.log
tail -25 gpurun_out/r2j_tests.log
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_mesh.py -x -q -k "small_and_odd or empty or overflow" > gpurun_out/r2j_memcheck.log 2>&1
tail -4 gpurun_out/r2j_memcheck.log; grep -c "Invalid\|out of bounds" gpurun_out/r2j_memcheck.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-train > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err
python -c "import json; d=json.loads(open('gpurun_out/r2j_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e'], d['e2e_all_logits'])"
tail -5 gpurun_out/r2j_bench.err
