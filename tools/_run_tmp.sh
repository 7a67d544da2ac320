set -x
mkdir -p gpurun_out
RFD_ONET_WAKE=1 timeout 900 python -m pytest tests/test_gpu_decoder.py -m gpu -x -q 2>&1 | tail -4
for rep in 1 2 3; do for s in 0 1; do
RFD_ONET_WAKE=$s timeout 300 python tools/prof_decoder.py 256 5 fp16 2 2>&1 | grep decode | sed "s/^/wake=$s /"
RFD_ONET_WAKE=$s timeout 300 python tools/prof_decoder.py 1024 3 fp16 2 2>&1 | grep decode | sed "s/^/wake=$s /"
done; done 2>&1 | grep -v "^+" | tee gpurun_out/r3c_wake_ab.log
