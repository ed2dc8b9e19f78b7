timeout 600 python -m pytest tests/test_gjk_gpu.py tests/test_pipeline_gpu.py -x -q 2>&1 | grep -E "^E|passed|failed" | head
timeout 300 python scripts/r02_dev.py pipe 2>&1 | grep -E "gjk|C5 shapes"
timeout 300 python scripts/gjk_per_type.py sphere,ellipsoid,capsule,cylinder,box 2>&1 | tail -2
