set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -70 > gpurun_out/r2d_tests.log
tail -12 gpurun_out/r2d_tests.log
timeout 300 python tools/prof_qg.py 4 30 2>&1 | tee gpurun_out/r2d_qg.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err
cat gpurun_out/r2d_bench.json; tail -5 gpurun_out/r2d_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"query_and_group|grid_build|transpose_features" -c 12 -o gpurun_out/r2d_qg python tools/prof_qg.py 4 1 > gpurun_out/r2d_ncu.log 2>&1
tail -3 gpurun_out/r2d_ncu.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r2d_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2d_ncu_bench.log 2>&1
