python bench.py > gpurun_out/bench_16o.json 2> gpurun_out/bench_16o.err; python -c "
import json; d=json.load(open('gpurun_out/bench_16o.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['traffic'], {k:v['ms_per_step'] for k,v in d['phases'].items() if v['ms_per_step']>0})"; tail -2 gpurun_out/bench_16o.err
