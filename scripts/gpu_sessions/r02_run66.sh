timeout 300 python -m pytest tests/test_pipeline_gpu.py -x -q 2>&1 | grep -E "^E|passed|failed" | head -3
timeout 300 python bench.py --only pipeline --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys, json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['stage_ms_rank0'])"
