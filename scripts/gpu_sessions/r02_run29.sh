python -m pytest tests/test_epa_gpu.py tests/test_gjk_gpu.py tests/test_pipeline_gpu.py -x -q 2>&1 | tail -3
python scripts/epa_thread_dev.py c5 types 2>&1 | tail -7
N=300000 ncu --set full --import-source on --clock-control none -k regex:k_epa_thread -c 1 -o gpurun_out/r02_epa_thread_v3 python scripts/epa_thread_dev.py c5 > gpurun_out/ncu_thread.log 2>&1
