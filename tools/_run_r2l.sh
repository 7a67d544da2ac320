set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_decoder.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2l_tests.log
tail -5 gpurun_out/r2l_tests.log
for prec in fp16 bf16 fp16x3; do timeout 300 python tools/prof_decoder.py 256 3 $prec 2; done 2>&1 | grep decode | tee gpurun_out/r2l_decoder_timing.log
timeout 300 python tools/prof_decoder.py 1024 3 fp16 2 2>&1 | grep decode | tee -a gpurun_out/r2l_decoder_timing.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-train --no-cpu-baseline > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err
python -c "import json; d=json.loads(open('gpurun_out/r2l_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline']['frac'], d['e2e']['value'], d['clocks'])"
