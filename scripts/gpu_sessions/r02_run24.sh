N=200000 ncu --set full --import-source on --clock-control none -k regex:k_epa_thread -c 1 -o gpurun_out/r02_epa_thread_v1 python scripts/epa_thread_dev.py c5 > gpurun_out/ncu_thread.log 2>&1
tail -3 gpurun_out/ncu_thread.log
