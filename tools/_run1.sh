set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_decoder.py -x -q -s 2>&1 | tail -40 > gpurun_out/r2a_dec_tests.log
cat gpurun_out/r2a_dec_tests.log | tail -15
for cl in 1 2; do for prec in fp16 bf16 fp16x3; do timeout 300 python tools/prof_decoder.py 256 3 $prec $cl; done; done 2>&1 | tee gpurun_out/r2a_prof.log
for cl in 1 2; do timeout 300 python tools/prof_decoder.py 1024 3 fp16 $cl; done 2>&1 | tee -a gpurun_out/r2a_prof.log
timeout 300 python -m pytest tests/test_gpu_ops.py -x -q -k abi 2>&1 | tail -3
