python bench.py --workload 12o --ansatz kupccgsd --batch 64 --steps 3 > gpurun_out/bench_12o_kup_b64.json 2>gpurun_out/err.txt; tail -2 gpurun_out/err.txt
python bench.py --workload 12o --ansatz kupccgsd --steps 5 --no-cpu-baseline > gpurun_out/bench_12o_kup.json 2>gpurun_out/err.txt; tail -2 gpurun_out/err.txt
python bench.py --workload 16o --ansatz puccd --batch 256 --steps 3 > gpurun_out/bench_16o_puccd_b256.json 2>gpurun_out/err.txt; tail -2 gpurun_out/err.txt
for f in 12o_kup_b64 12o_kup 16o_puccd_b256; do python -c "import json; d=json.load(open('gpurun_out/bench_$f.json')); print('$f', 'evals/s %.1f e2e %.1f ms/step %.2f launches %d'%(d['value'], d['e2e']['value'], d['ms_per_step'], d['gpu_launches']), d.get('cpu_baseline',{}).get('value'))"; done
