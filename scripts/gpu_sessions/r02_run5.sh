set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --no-extra > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err || { tail -30 gpurun_out/bench_c1.err; exit 1; }
python -c "
import json; d=json.load(open('gpurun_out/bench_c1.json'))
print(d['value'], json.dumps(d['e2e']))
"
python bench.py --only broad 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read())
for k in ('dense','constant_density'): print(k, {x:d[k][x] for x in ('build_ms','query_ms','unique_overlap_pairs','overlap_pairs_per_s','ordered_form_query_ms')}, d[k]['roofline_query']['frac_compulsory'], d[k]['parity_vs_brute_force'])
"
python scripts/r02_dev.py epa pipe 2>&1 | grep -E "epa|EPA|C5"
