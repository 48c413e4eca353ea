python -m pytest tests -x -q -m gpu 2>&1 | tail -2
python bench.py > gpurun_out/bench_16o.json 2> gpurun_out/bench_16o.err; tail -c 3500 gpurun_out/bench_16o.json; tail -3 gpurun_out/bench_16o.err
python bench.py --workload 12o --no-cpu-baseline > gpurun_out/bench_12o.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_12o.json')); print('12o', d['value'], d['e2e']['value'], d['ms_per_step'])"
python bench.py --workload 8o --no-cpu-baseline --steps 20 > gpurun_out/bench_8o.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_8o.json')); print('8o', d['value'], d['e2e']['value'], d['ms_per_step'])"
python bench.py --workload 4o --no-cpu-baseline --steps 20 > gpurun_out/bench_4o.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_4o.json')); print('4o', d['value'], d['e2e']['value'], d['ms_per_step'])"
