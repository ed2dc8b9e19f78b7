time timeout 900 python bench.py > gpurun_out/bench_full_n1_v5.json 2> gpurun_out/bench_full_n1_v5.err || tail -30 gpurun_out/bench_full_n1_v5.err
python -c "
import json; d=json.load(open('gpurun_out/bench_full_n1_v5.json'))
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['value_including_host_packing'], d['gpu_launches'], d['clocks'])
for k in ('epa','self_collision','pipeline','six_type_mix','libccd','hydroelastic'):
    if k in d: print(k, d[k]['value'], d[k].get('parity_on_cpu_sample'), d[k].get('stage_ms_rank0'))
print(d['parity_on_cpu_sample'], d['cpu_baseline'])
"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | cut -c1-600
