set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mesh.py tests/test_gpu_modules.py -m gpu -x -q -k "scene_generation or inplace_weight" 2>&1 | tail -8
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-train 2>gpurun_out/r3d_err.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value']); print(d['full_generation']); print(d['skip_propagation']['generate_ms'])" || tail -20 gpurun_out/r3d_err.log
