set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_modules.py -m gpu -x -q -s -k "skip_propagation or stn_group or chain" 2>&1 | tail -30
