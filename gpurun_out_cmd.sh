set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
CMD="python bench.py --workload 12o --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_12o.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 10929 -c 3700 --csv --log-file gpurun_out/launches_12o.csv $CMD > gpurun_out/ncu_12o.log 2>&1
tail -2 gpurun_out/ncu_12o.log
CMD2="python bench.py --workload 16o --steps 1 --warmup 3 --no-cpu-baseline"
$CMD2 > gpurun_out/plain_16o.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:reverse_kernel -s 2000 -c 2 -o gpurun_out/prof_rev_16o $CMD2 > gpurun_out/ncu_16o.log 2>&1
tail -2 gpurun_out/ncu_16o.log
ls -la gpurun_out
