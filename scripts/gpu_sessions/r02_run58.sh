timeout 600 python -m pytest tests/test_self_collision_gpu.py -x -q 2>&1 | grep -E "^E|passed|failed" | head
timeout 300 python bench.py --only selfcollision --no-cpu-baseline 2>&1 | tail -1 | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_c4_v2.csv python bench.py --only selfcollision --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b_c4.log 2>&1
