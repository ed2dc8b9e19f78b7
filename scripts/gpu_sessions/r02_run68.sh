timeout 600 python -m pytest tests/test_gjk_gpu.py -x -q -k "wire or stream" 2>&1 | grep -E "^E|passed|failed" | head
time timeout 600 python bench.py --no-extra --no-cpu-baseline > gpurun_out/bench_n1_noextra.json 2> gpurun_out/bench_n1_noextra.err || tail -30 gpurun_out/bench_n1_noextra.err
python -c "
import json; d=json.load(open('gpurun_out/bench_n1_noextra.json'))
print(d['value'], d['ms_per_step']); print(json.dumps(d['e2e']))
"
