N=300000 ncu --set full --import-source on --clock-control none -k regex:k_epa -c 5 -o gpurun_out/r02_epa_thread_v2 python scripts/epa_thread_dev.py c5 > gpurun_out/ncu_thread.log 2>&1
tail -3 gpurun_out/ncu_thread.log
