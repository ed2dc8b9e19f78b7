python -m pytest tests/test_epa_gpu.py tests/test_gjk_gpu.py tests/test_pipeline_gpu.py tests/test_self_collision_gpu.py -x -q 2>&1 | grep -E "^E|passed|failed" | head -20
python scripts/r02_dev.py epa pipe 2>&1 | grep -E "epa|EPA|C5 shapes|gjk"
N=300000 ncu --set full --import-source on --clock-control none -k regex:k_epa_thread -c 1 -o gpurun_out/r02_epa_thread_v4 python scripts/epa_thread_dev.py c5 > gpurun_out/ncu_thread.log 2>&1
