python scripts/epa_thread_dev.py 2>&1 | tail -12
N=200000 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_epa --csv --log-file gpurun_out/epa_thread_launches.csv python scripts/epa_thread_dev.py c5 > /dev/null 2>&1
grep -E "k_epa" gpurun_out/epa_thread_launches.csv | awk -F'","' '{print $5, $NF}' | cut -c1-120 | tail -16
