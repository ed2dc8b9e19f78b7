set -x
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --no-extra --no-cpu-baseline --steps 2 --warmup 1 > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_epa -s 4 -c 1 -o gpurun_out/r02_epa_c5 python scripts/r02_dev.py pipe > gpurun_out/ncu_epa_c5.log 2>&1
tail -3 gpurun_out/ncu_epa_c5.log
