set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python scripts/r02_dev.py bvh epa pipe 2>&1 | tee gpurun_out/r02_dev_v1.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_epa$ -c 1 -o gpurun_out/r02_epa_v3 python scripts/r02_dev.py epa > gpurun_out/ncu_epa.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_overlap_append -s 3 -c 1 -o gpurun_out/r02_overlap_self python scripts/r02_dev.py bvh > gpurun_out/ncu_bvh.log 2>&1
ls -la gpurun_out
