set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_modules.py tests/test_gpu_dropin.py -m gpu -x -q 2>&1 | tail -30 > gpurun_out/r2i_tests.log
tail -5 gpurun_out/r2i_tests.log
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_ops.py -x -q -k "fused_query_and_group or grid_path or ball_query" > gpurun_out/r2i_memcheck.log 2>&1
grep -n "=========" gpurun_out/r2i_memcheck.log | head -30
timeout 300 python tools/prof_qg.py 4 30 2>&1 | tee gpurun_out/r2i_qg.log
