set -x
python -m pytest tests/test_epa_gpu.py tests/test_pipeline_gpu.py tests/test_gjk_gpu.py -m gpu -x -q 2>&1 | tail -5
python scripts/r02_dev.py epa pipe 2>&1 | grep -E "epa|EPA|C5 shapes" | tee gpurun_out/r02_epa_ids.txt
D3D_EPA_EXACT_EDGES=1 python scripts/r02_dev.py epa pipe 2>&1 | grep -E "epa|EPA|C5 shapes" | tee -a gpurun_out/r02_epa_ids.txt
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:^k_epa$" -s 1 -c 1 -o gpurun_out/r02_epa_c5 python scripts/r02_dev.py pipe > gpurun_out/ncu_epa_c5.log 2>&1
tail -2 gpurun_out/ncu_epa_c5.log
