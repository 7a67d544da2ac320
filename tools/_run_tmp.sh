set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_modules.py -m gpu -x -q -k "cuda_graph" 2>&1 | tail -12
for a in "" "--graph-detection"; do
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-train $a 2>gpurun_out/r2w_err.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$a', '| value', round(d['value'],2), 'ms', round(d['ms_per_step'],2), '| e2e', round(d['e2e']['value'],2), 'logits', round(d['e2e_all_logits']['value'],2), 'bits', round(d['e2e_occupancy_bits']['value'],2), '| dec', round(d['roofline']['ms_per_launch'],2))" || tail -5 gpurun_out/r2w_err.log
done 2>&1 | grep -v "^+" | tee gpurun_out/r2w_variants.log
