set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_modules.py tests/test_gpu_dropin.py -m gpu -x -q 2>&1 | tail -30 > gpurun_out/r2h_tests.log
tail -8 gpurun_out/r2h_tests.log
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_ops.py -x -q -k "fused_query_and_group or grid_path or ball_query" 2>&1 | tail -12 > gpurun_out/r2h_memcheck.log
tail -5 gpurun_out/r2h_memcheck.log
timeout 300 python tools/prof_qg.py 4 30 2>&1 | tee gpurun_out/r2h_qg.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"query_and_group|grid_build|transpose_features" -c 12 -o gpurun_out/r2h_qg python tools/prof_qg.py 4 1 > gpurun_out/r2h_ncu.log 2>&1
tail -3 gpurun_out/r2h_ncu.log
