set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mesh.py -m gpu -x -q -s 2>&1 | tail -30 > gpurun_out/r2k_tests.log
tail -8 gpurun_out/r2k_tests.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-train > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err
python -c "import json; d=json.loads(open('gpurun_out/r2k_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e'], d['e2e_all_logits'], d['e2e_occupancy_bits'], d['cpu_baseline'])"
tail -5 gpurun_out/r2k_bench.err
