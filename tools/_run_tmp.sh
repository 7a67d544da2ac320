set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_mesh.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/prof_mesh.py 256 2>&1 | tail -2 | tee gpurun_out/r2r_mesh.log
