timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_c4.csv python bench.py --only selfcollision --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b_c4.log 2>&1
tail -1 gpurun_out/b_c4.log | cut -c1-400
