set -x
python -m pytest tests/test_broad_phase_gpu.py tests/test_pipeline_gpu.py tests/test_hydroelastic_gpu.py tests/test_self_collision_gpu.py -m gpu -x -q 2>&1 | tail -8
python scripts/r02_dev.py bvh 2>&1 | tee gpurun_out/r02_dev_v3_tiles.txt
python bench.py --only broad 2>&1 | tail -1 > gpurun_out/bench_broad_tiles.json
python -c "
import json; d=json.load(open('gpurun_out/bench_broad_tiles.json'))
for k in ('dense','constant_density'): print(k, {x:d[k][x] for x in ('build_ms','query_ms','unique_overlap_pairs','overlap_pairs_per_s','ordered_form_query_ms','node_visits')}, d[k]['roofline_query'], d[k]['parity_vs_brute_force'])
"
