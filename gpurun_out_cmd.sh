timeout 900 python bench.py --workload 18o --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_18o.json 2> gpurun_out/bench_18o.err; tail -c 1800 gpurun_out/bench_18o.json; tail -3 gpurun_out/bench_18o.err
nvidia-smi --query-gpu=memory.used --format=csv
