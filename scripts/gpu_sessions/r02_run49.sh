time timeout 900 python bench.py --no-extra > gpurun_out/bench_n1_noextra.json 2> gpurun_out/bench_n1_noextra.err || tail -30 gpurun_out/bench_n1_noextra.err
python -c "
import json; d=json.load(open('gpurun_out/bench_n1_noextra.json'))
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['gpu_launches']); print(json.dumps(d['e2e']))
"
