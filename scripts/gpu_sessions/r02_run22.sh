set -x
python -m pytest tests/test_epa_gpu.py tests/test_gjk_gpu.py tests/test_pipeline_gpu.py -x -q 2>&1 | tail -15
python scripts/r02_dev.py epa pipe 2>&1 | grep -E "epa|EPA|C5 shapes|gjk"
D3D_EPA_KERNEL=warp python scripts/r02_dev.py epa 2>&1 | grep -E "epa|EPA"
