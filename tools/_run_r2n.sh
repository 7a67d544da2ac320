set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_decoder.py -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2n_tests.log
tail -25 gpurun_out/r2n_tests.log
timeout 300 python tools/trace_decoder.py 2>&1 | tee gpurun_out/r2n_trace.log | tail -8
for prec in fp16 bf16 fp16x3; do timeout 300 python tools/prof_decoder.py 256 5 $prec 2; done 2>&1 | grep decode | tee gpurun_out/r2n_decoder_timing.log
timeout 300 python tools/prof_decoder.py 256 5 fp16 1 2>&1 | grep decode | tee -a gpurun_out/r2n_decoder_timing.log
timeout 300 python tools/prof_decoder.py 1024 3 fp16 2 2>&1 | grep decode | tee -a gpurun_out/r2n_decoder_timing.log
