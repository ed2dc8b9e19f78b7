timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_c1.csv python bench.py --steps 2 --warmup 1 --no-extra --no-cpu-baseline > gpurun_out/b_c1.log 2>&1
tail -1 gpurun_out/b_c1.log | cut -c1-200
