timeout 600 python -m pytest tests/test_self_collision_gpu.py tests/test_pipeline_gpu.py -x -q 2>&1 | grep -E "^E|passed|failed" | head
timeout 300 python bench.py --only selfcollision 2>&1 | tail -1 | cut -c1-900
