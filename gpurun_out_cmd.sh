CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_16o.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 2343 -c 781 --csv --log-file gpurun_out/launches_16o.csv $CMD > gpurun_out/ncu_l.log 2>&1
tail -1 gpurun_out/ncu_l.log | cut -c1-200
$CMD > gpurun_out/plain_16o.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:batch_kernel -s 1288 -c 6 -o gpurun_out/prof_final_batch_16o $CMD > gpurun_out/ncu_f.log 2>&1
tail -1 gpurun_out/ncu_f.log | cut -c1-200
