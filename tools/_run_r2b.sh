set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2b_gputests.log
tail -5 gpurun_out/r2b_gputests.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
cat gpurun_out/r2b_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2b_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_ncu_bench.log 2>&1
tail -2 gpurun_out/r2b_ncu_bench.log | cut -c1-300
