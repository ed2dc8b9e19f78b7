set -x
python -m pytest tests/test_epa_gpu.py tests/test_pipeline_gpu.py -m gpu -x -q 2>&1 | tail -3
python scripts/epa_retry_rate.py
python scripts/r02_dev.py epa pipe 2>&1 | grep -E "epa|EPA|C5 shapes" | tee gpurun_out/r02_epa_ids.txt
