set -x
python scripts/epa_retry_rate.py 2>&1 | tail -2
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:^k_epa$" -s 0 -c 1 -o gpurun_out/r02_epa_c5_ids python scripts/epa_retry_rate.py > gpurun_out/ncu_epa_c5.log 2>&1
tail -2 gpurun_out/ncu_epa_c5.log
D3D_EPA_EXACT_EDGES=1 timeout 900 ncu --set full --clock-control none --import-source on -k "regex:^k_epa$" -s 0 -c 1 -o gpurun_out/r02_epa_c5_exact python scripts/epa_retry_rate.py > gpurun_out/ncu_epa_c5b.log 2>&1
tail -2 gpurun_out/ncu_epa_c5b.log
