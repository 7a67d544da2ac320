set -x
mkdir -p gpurun_out
for f in 1 0; do
RFD_ONET_FUSE_E0=$f timeout 300 python tools/trace_decoder.py 2>&1 | tee gpurun_out/r2m_trace_fuse$f.log | head -16
for i in 1 2; do RFD_ONET_FUSE_E0=$f timeout 300 python tools/prof_decoder.py 256 5 fp16 2 2>&1 | grep decode; done
RFD_ONET_FUSE_E0=$f timeout 300 python tools/prof_decoder.py 1024 3 fp16 2 2>&1 | grep decode
done 2>&1 | tee gpurun_out/r2m_log.txt
