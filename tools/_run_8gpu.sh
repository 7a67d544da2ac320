set -x
mkdir -p gpurun_out
nvidia-smi -L | head -8
for n in 8 4; do
NCCL_DEBUG=INFO timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2961$n bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/r2s_bench_${n}gpu.json 2> gpurun_out/r2s_bench_${n}gpu.err
python -c "import json,sys; d=json.loads(open('gpurun_out/r2s_bench_${n}gpu.json').read().strip().splitlines()[0]); print(d['n_gpus'], d['value'], d['e2e']['value']); t=d['train']; print({k:t.get(k) for k in ['ms_per_step','scenes_per_s','allreduce_us','bus_gbs','exposed_us','overlap_frac','error']})"
grep -i "NVLS\|nranks" gpurun_out/r2s_bench_${n}gpu.err | head -4
grep -v "NCCL INFO" gpurun_out/r2s_bench_${n}gpu.err | tail -5
done
