set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/r2g_tests.log
tail -8 gpurun_out/r2g_tests.log
NCCL_DEBUG=INFO timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2g_bench_2gpu.json 2> gpurun_out/r2g_bench_2gpu.err
python -c "import json; d=json.loads(open('gpurun_out/r2g_bench_2gpu.json').read().strip().splitlines()[-1]); print(json.dumps(d['train'],indent=1)); print(d['value'], d['e2e'])"
grep -i "NVLS\|nranks\|comm 0x" gpurun_out/r2g_bench_2gpu.err | head -12
tail -5 gpurun_out/r2g_bench_2gpu.err
