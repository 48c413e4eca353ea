python bench.py --workload 4o --batch 4096 --steps 3 --no-cpu-baseline > gpurun_out/bench_4o_b4096.json 2>gpurun_out/err.txt; tail -2 gpurun_out/err.txt
python bench.py --workload 8o --batch 1024 --steps 3 > gpurun_out/bench_8o_b1024.json 2>gpurun_out/err.txt; tail -2 gpurun_out/err.txt
for f in 4o_b4096 8o_b1024; do python -c "import json; d=json.load(open('gpurun_out/bench_$f.json')); print('$f', 'evals/s %.1f e2e %.1f ms/step %.2f launches %d'%(d['value'], d['e2e']['value'], d['ms_per_step'], d['gpu_launches']), d.get('cpu_baseline',{}).get('value'))"; done
