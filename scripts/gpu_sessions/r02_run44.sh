time timeout 900 python bench.py > gpurun_out/bench_full_n1_v4.json 2> gpurun_out/bench_full_n1_v4.err || tail -30 gpurun_out/bench_full_n1_v4.err
python -c "
import json; d=json.load(open('gpurun_out/bench_full_n1_v4.json'))
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['gpu_launches'], d['clocks'])
bp=d['broad_phase']
for k in ('dense','constant_density'): print(k, {x:bp[k][x] for x in ('build_ms','query_ms','overlap_pairs_per_s')}, bp[k]['roofline_query']['frac_compulsory'], bp[k]['parity_vs_brute_force']['sets_equal'])
for k in ('epa','self_collision','pipeline','six_type_mix','libccd','hydroelastic'):
    if k in d: print(k, d[k]['value'], d[k].get('parity_on_cpu_sample'), d[k].get('stage_ms_rank0'))
"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_pipeline.csv python bench.py --only pipeline --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b_pipe.log 2>&1
tail -2 gpurun_out/b_pipe.log | cut -c1-200
