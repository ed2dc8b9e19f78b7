set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
export D3D_SAN_COLLIDERS=200 D3D_SAN_PAIRS=600
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python scripts/sanitize.py > gpurun_out/r02_memcheck.txt 2>&1; echo "memcheck rc=$?" >> gpurun_out/r02_memcheck.txt
tail -4 gpurun_out/r02_memcheck.txt
export D3D_SAN_COLLIDERS=80 D3D_SAN_PAIRS=150
timeout 1500 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 30 python scripts/sanitize.py > gpurun_out/r02_racecheck.txt 2>&1; echo "racecheck rc=$?" >> gpurun_out/r02_racecheck.txt
tail -6 gpurun_out/r02_racecheck.txt
unset D3D_SAN_COLLIDERS D3D_SAN_PAIRS
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tile -s 2 -c 1 -o gpurun_out/r02_tile_self python scripts/r02_dev.py bvh > gpurun_out/ncu_tile.log 2>&1
time python bench.py > gpurun_out/bench_full_n1_v2.json 2> gpurun_out/bench_full_n1_v2.err || tail -30 gpurun_out/bench_full_n1_v2.err
python -c "
import json; d=json.load(open('gpurun_out/bench_full_n1_v2.json'))
print(d['value'], d['ms_per_step'], json.dumps(d['roofline'])[:400]); print(json.dumps(d['e2e']))
bp=d['broad_phase']
for k in ('dense','constant_density'): print(k, {x:bp[k][x] for x in ('build_ms','query_ms','overlap_pairs_per_s','ordered_form_query_ms')}, bp[k]['roofline_query']['frac_compulsory'], bp[k]['parity_vs_brute_force']['sets_equal'])
for k in ('epa','self_collision','pipeline','six_type_mix'): print(k, d[k]['value'], d[k].get('parity_on_cpu_sample'), d[k].get('stage_ms_rank0'))
print(d['parity_on_cpu_sample'], d['cpu_baseline'])
"
