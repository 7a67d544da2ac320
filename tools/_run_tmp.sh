set -x
mkdir -p gpurun_out
for nb in 2 1 2 1; do RFD_QG_NBUF=$nb timeout 300 python tools/prof_qg.py 4 30 2>&1 | sed "s/^/nbuf=$nb /"; done | tee gpurun_out/r2v_qg.log
RFD_QG_NBUF=1 timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "fused or group" 2>&1 | tail -2
