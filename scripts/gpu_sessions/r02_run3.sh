set -x
time python bench.py > gpurun_out/bench_full_n1.json 2> gpurun_out/bench_full_n1.err || { tail -30 gpurun_out/bench_full_n1.err; exit 1; }
python -c "
import json; d=json.load(open('gpurun_out/bench_full_n1.json'))
print({k:(v if not isinstance(v,dict) else '...') for k,v in d.items()})
for k in ('e2e','roofline','broad_phase','epa','self_collision','pipeline','six_type_mix','cpu_baseline','parity_on_cpu_sample'): print(k, json.dumps(d[k])[:2600])
"
