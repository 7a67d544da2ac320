set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/r2e_tests.log
tail -8 gpurun_out/r2e_tests.log
timeout 300 python tools/prof_qg.py 4 30 2>&1 | tee gpurun_out/r2e_qg.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err
cat gpurun_out/r2e_bench.json; tail -5 gpurun_out/r2e_bench.err
