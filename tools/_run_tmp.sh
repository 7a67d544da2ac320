set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_modules.py -m gpu -x -q -k "chain_ex" 2>&1 | tail -6
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_modules.py -x -q -k "chain_ex or stn_group_vs_reference" > gpurun_out/r3a_memcheck.log 2>&1
tail -3 gpurun_out/r3a_memcheck.log; grep -c "Invalid\|out of bounds\|Misaligned" gpurun_out/r3a_memcheck.log
timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_mesh.py -x -q -k "small_and_odd" > gpurun_out/r3a_racecheck.log 2>&1
tail -3 gpurun_out/r3a_racecheck.log
