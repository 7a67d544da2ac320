set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_ops.py tests/test_gpu_modules.py -x -q -s 2>&1 | tail -60 > gpurun_out/r2c_tests.log
tail -12 gpurun_out/r2c_tests.log
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_ops.py -x -q -k "fused_query_and_group or grid_path" 2>&1 | tail -25 > gpurun_out/r2c_memcheck.log
tail -6 gpurun_out/r2c_memcheck.log
timeout 300 python tools/prof_qg.py 4 30 2>&1 | tee gpurun_out/r2c_qg.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err
cat gpurun_out/r2c_bench.json; tail -5 gpurun_out/r2c_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"query_and_group|grid_build|transpose_features" -c 12 -o gpurun_out/r2c_qg python tools/prof_qg.py 4 1 > gpurun_out/r2c_ncu.log 2>&1
tail -3 gpurun_out/r2c_ncu.log
