timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | grep -E "^E|passed|failed" | head -5
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
time timeout 900 python bench.py > gpurun_out/bench_full_n1_v7.json 2> gpurun_out/bench_full_n1_v7.err || tail -30 gpurun_out/bench_full_n1_v7.err
python -c "
import json; d=json.load(open('gpurun_out/bench_full_n1_v7.json'))
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['value_including_host_packing'], d['gpu_launches'], d['clocks'])
for k in ('epa','self_collision','pipeline','six_type_mix','libccd','hydroelastic'):
    if k in d: print(k, d[k]['value'], d[k].get('parity_on_cpu_sample'), d[k].get('stage_ms_rank0'))
print(d['parity_on_cpu_sample'], d['cpu_baseline'])
"
